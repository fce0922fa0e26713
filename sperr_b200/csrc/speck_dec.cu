// Integer SPECK decoders: tree policies of the 3D coefficient coder and the 1D outlier coder for the
// generic engine (speck_dec.cuh), the fast engine for power-of-two trees (speck_dec_fast.cuh) and
// the host driver that runs every stream of a batch side by side.
//   set partitioning      /root/reference/src/SPECK3D_INT.cpp:214-326 (m_partition_S_XYZ and friends)
//   initial sets          /root/reference/src/SPECK3D_INT.cpp:22-97
//   decoder significance  /root/reference/src/SPECK3D_INT_DEC.cpp:8-49
//   SPECK1D_INT / _DEC    /root/reference/src/SPECK1D_INT.cpp:18-56, src/SPECK1D_INT_DEC.cpp:12-125
#include "speck_dec_fast.cuh"
#include "tree3d.cuh"

namespace sperr_b200 {

struct DecTree3D {
  static constexpr int kKind = 0;
  static constexpr bool kHasI = false;
  template <class D>
  static __device__ __forceinline__ node_t iset(const D&, const DecChunk&, unsigned) { return 0; }
  struct Data {
    const ShapeDev* shapes;
  };

  static __device__ __forceinline__ int num_roots(const Data& t, const DecChunk& d, unsigned)
  {
    return t.shapes[d.shape].h->nroots;
  }

  static __device__ __forceinline__ void root(const Data& t, const DecChunk& d, unsigned, int r,
                                              node_t& nd, int& lis)
  {
    const RootDesc& rd = t.shapes[d.shape].h->roots[r];
    nd = make_node(rd.level, rd.ix, rd.iy, rd.iz);
    lis = rd.lis;
  }

  // Children in the reference's order (x fastest), empty ones dropped.
  static __device__ __forceinline__ int children(const Data& t, const DecChunk& d, unsigned,
                                                 node_t nd, int, DChild* out)
  {
    const ShapeDev s = t.shapes[d.shape];
    NodeGeom g;
    node_geom(s, nd, g);
    const ShapeHeader* h = s.h;
    int k = 0;
    if (g.Lc == h->leaf_level) {   // every child is a single coefficient
      for (unsigned cz = 0; cz < g.nzc; cz++)
        for (unsigned cy = 0; cy < g.nyc; cy++)
          for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
            DChild& r = out[k];
            r.pixel = 1;
            r.id = 0;
            r.lis = 0;
            r.idx = ((unsigned long long)(g.z0 + cz) * h->ny + (g.y0 + cy)) * h->nx + (g.x0 + cx);
          }
      return k;
    }
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
          const unsigned jx = g.x0 + cx, jy = g.y0 + cy, jz = g.z0 + cz;
          DChild& r = out[k];
          unsigned lx, ly, lz;
          node_len(s, g.Lc, jx, jy, jz, lx, ly, lz);
          if (lx * ly * lz == 1) {
            r.pixel = 1;
            r.id = 0;
            r.lis = 0;
            r.idx = node_raster(s, g.Lc, jx, jy, jz);
          }
          else {
            r.pixel = 0;
            r.id = make_node(g.Lc, jx, jy, jz);
            r.idx = 0;
            r.lis = int(node_lis(s, g.Lc, jx, jy, jz));
          }
        }
    return k;
  }
};


// 2D slices: quadtree in reverse raster order plus the set I (SPECK2D_INT,
// /root/reference/src/SPECK2D_INT.cpp:60-213, src/SPECK2D_INT_DEC.cpp:5-52). I at part_level l is
// node (level 0xFF, ix = l); its list index is -1 (it lives outside the lists).
struct DecTree2D {
  static constexpr int kKind = 2;
  static constexpr bool kHasI = true;
  typedef DecTree3D::Data Data;
  static constexpr int kINode = 0xFF;

  static __device__ __forceinline__ node_t iset(const Data& t, const DecChunk& d, unsigned)
  {
    const int l = t.shapes[d.shape].h->nxf2d;
    return l > 0 ? make_node(kINode, unsigned(l), 0, 0) : 0;
  }
  static __device__ __forceinline__ int num_roots(const Data&, const DecChunk&, unsigned) { return 1; }
  static __device__ __forceinline__ void root(const Data& t, const DecChunk& d, unsigned, int,
                                              node_t& nd, int& lis)
  {
    const ShapeHeader* h = t.shapes[d.shape].h;
    nd = make_node(h->lv2d[h->nxf2d], 0, 0, 0);
    lis = h->nxf2d;
  }
  static __device__ __forceinline__ void fill(const ShapeDev& s, int L, unsigned jx, unsigned jy,
                                              DChild& r)
  {
    unsigned lx, ly, lz;
    node_len(s, L, jx, jy, 0, lx, ly, lz);
    if (lx * ly == 1) {
      r.pixel = 1;
      r.id = 0;
      r.lis = 0;
      r.idx = node_raster(s, L, jx, jy, 0);
    }
    else {
      r.pixel = 0;
      r.id = make_node(L, jx, jy, 0);
      r.idx = 0;
      r.lis = s.h->lv[L].j;   // part_level = chain position
    }
  }
  static __device__ __forceinline__ int children(const Data& t, const DecChunk& d, unsigned,
                                                 node_t nd, int, DChild* out)
  {
    const ShapeDev s = t.shapes[d.shape];
    const ShapeHeader* h = s.h;
    if (node_level(nd) == kINode) {
      const int l = int(node_ix(nd));
      const int L = h->lv2d[l];
      fill(s, L, 1, 1, out[0]);   // BR, TR, BL (src/SPECK2D_INT.cpp:150-185)
      fill(s, L, 1, 0, out[1]);
      fill(s, L, 0, 1, out[2]);
      if (l == 1)
        return 3 | 0x100;
      out[3].pixel = 0;
      out[3].id = make_node(kINode, unsigned(l - 1), 0, 0);
      out[3].idx = 0;
      out[3].lis = -1;
      return 4;
    }
    NodeGeom g;
    node_geom(s, nd, g);
    int k = 0;
    for (int cy = int(g.nyc) - 1; cy >= 0; cy--)
      for (int cx = int(g.nxc) - 1; cx >= 0; cx--, k++)
        fill(s, g.Lc, g.x0 + cx, g.y0 + cy, out[k]);
    return k;
  }
};

// node = start | len << 32 ; list index = depth of the set (the two initial halves are depth 1)
struct DecTree1D {
  static constexpr int kKind = 1;
  static constexpr bool kHasI = false;
  struct Data {
    int unused;
  };
  static __device__ __forceinline__ node_t iset(const Data&, const DecChunk&, unsigned) { return 0; }
  static __device__ __forceinline__ int num_roots(const Data&, const DecChunk&, unsigned) { return 2; }
  static __device__ __forceinline__ void root(const Data&, const DecChunk& d, unsigned, int r,
                                              node_t& nd, int& lis)
  {
    const unsigned long long n = d.n, first = n - n / 2;
    nd = r == 0 ? (first << 32) : (first | ((n / 2) << 32));
    lis = 1;
  }
  static __device__ __forceinline__ int children(const Data&, const DecChunk&, unsigned, node_t nd,
                                                 int lis, DChild* out)
  {
    const unsigned long long start = nd & 0xffffffffull, len = nd >> 32;
    const unsigned long long l0 = len - len / 2, l1 = len / 2;
    out[0].pixel = l0 == 1;
    out[0].idx = start;
    out[0].id = start | (l0 << 32);
    out[0].lis = lis + 1;
    out[1].pixel = l1 == 1;
    out[1].idx = start + l0;
    out[1].id = (start + l0) | (l1 << 32);
    out[1].lis = lis + 1;
    return 2;
  }
};


// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------

// out[w] = little-endian word w of the byte string src[0, len), zero beyond it
static __global__ void k_stage_bits(const DecChunk* chunks, const unsigned char* const* srcs,
                                    const unsigned long long* lens, const unsigned long long* words)
{
  const unsigned c = blockIdx.y;
  const unsigned char* src = srcs[c];
  const unsigned long long len = lens[c], nw = words[c];
  uint32_t* out = const_cast<uint32_t*>(chunks[c].bits);
  const unsigned long long avail = chunks[c].avail;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nw;
       w += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) {
      const unsigned long long i = w * 4 + b;
      if (i < len)
        v |= uint32_t(src[i]) << (8 * b);
    }
    if (w * 32 >= avail)   // bits the stream does not really hold read as zero
      v = 0;
    else if (avail - w * 32 < 32)
      v &= (1u << (avail - w * 32)) - 1u;
    out[w] = v;
  }
}

// Decodes the sorting passes of every job (3D coefficient streams and 1D outlier streams run side
// by side: one CTA each); w.h / w.dchunks hold the decoder state afterwards.
static size_t lipsum_words(size_t mask_words_of_chunk)
{
  return ((mask_words_of_chunk / 128 + 1 + 31) / 32 + 3) & ~size_t(3);   // keeps the pieces 16-byte aligned
}

void speck_decode(DecWork& w, const std::vector<DecJob>& jobs, const ShapeDev* d_shapes,
                  cudaStream_t st)
{
  const int nj = int(jobs.size());
  if (nj == 0)
    return;
  size_t mask_words = 0, lis_entries = 0, cnt_entries = 0, stage_words = 0, pl_bytes = 0;
  bool any_fast = false, any_slow3 = false, any_slow1 = false, any_slow2 = false;
  std::vector<size_t> mw(nj), sw(nj);
  w.max_n = 0;
  w.max_planes = 0;
  for (int c = 0; c < nj; c++) {
    const DecJob& j = jobs[c];
    mw[c] = j.skip ? 0 : ((size_t(j.n + 31) / 32 + 8) & ~size_t(3));   // 16-byte aligned pieces, zero padded
    sw[c] = j.skip ? 0 : (size_t(j.payload_bytes) / 4 + 4);
    mask_words += 3 * mw[c] + lipsum_words(mw[c]);
    pl_bytes += j.skip ? 0 : ((size_t(j.n) + 63) & ~size_t(63));
    lis_entries += j.skip ? 0 : j.lis_total;
    cnt_entries += j.skip ? 0 : size_t(j.nlis + 1);
    stage_words += sw[c];
    if (!j.skip) {
      (j.pow2 ? any_fast : (j.kind == 0 ? any_slow3 : (j.kind == 2 ? any_slow2 : any_slow1))) = true;
      w.max_n = std::max<size_t>(w.max_n, j.n);
      w.max_planes = std::max(w.max_planes, j.planes);
      if (j.planes > kMaxPlanes)
        throw std::runtime_error("SPECK stream with more than 64 bit-planes");
    }
  }
  w.masks.reserve(mask_words * 4 + 16);
  w.pl.reserve(pl_bytes + 16);
  w.lis.reserve(lis_entries * 8 + 16);
  w.lis_cnt.reserve(cnt_entries * 4 + 16);
  w.stage.reserve(stage_words * 4 + 16);
  rt::dset(w.masks.p, 0, mask_words * 4, st);
  rt::dset(w.pl.p, 0xFF, pl_bytes, st);
  w.h.assign(nj, DecChunk());
  std::vector<const unsigned char*> srcs(nj);
  std::vector<unsigned long long> lens(nj), words(nj);
  size_t om = 0, ol = 0, oc = 0, os = 0, op = 0, max_words = 1;
  for (int c = 0; c < nj; c++) {
    const DecJob& j = jobs[c];
    DecChunk& d = w.h[c];
    std::memset(&d, 0, sizeof(d));
    d.skip = j.skip ? 1 : 0;
    srcs[c] = j.d_payload;
    lens[c] = j.skip ? 0 : j.payload_bytes;
    words[c] = sw[c];
    if (j.skip)
      continue;
    d.n = j.n;
    d.shape = j.shape;
    d.kind = j.kind;
    d.planes = j.planes;
    d.avail = std::min<unsigned long long>(j.total_bits, j.payload_bytes * 8ull);
    uint32_t* m = w.masks.as<uint32_t>() + om;
    d.lip = m;
    d.sigarr = m + mw[c];
    d.signarr = m + 2 * mw[c];
    d.lipsum = m + 3 * mw[c];
    om += 3 * mw[c] + lipsum_words(mw[c]);
    d.pl = w.pl.as<uint8_t>() + op;
    op += (size_t(j.n) + 63) & ~size_t(63);
    d.lis = w.lis.as<node_t>() + ol;
    ol += j.lis_total;
    d.lis_off = j.d_lis_off;
    d.lis_cnt = w.lis_cnt.as<unsigned>() + oc;
    oc += size_t(j.nlis + 1);
    d.nlis = j.nlis;
    d.bits = w.stage.as<uint32_t>() + os;
    d.stage_words = sw[c];
    os += sw[c];
    d.pow2 = j.pow2;
    d.Dx = j.Dx; d.Dy = j.Dy; d.Dz = j.Dz;
    d.nx = j.nx; d.ny = j.ny;
    d.nroots = j.nroots;
    d.iset = node_t(j.nxf2d);   // fast 2D path: I's part_level (the generic kernel sets its own)
    for (int r = 0; r < j.nroots; r++)
      d.roots[r] = j.roots[r];
    max_words = std::max(max_words, sw[c]);
  }
  // Clusters: the chain phases of a power-of-two 3D / 2D stream are shared by R CTAs (R consecutive
  // windows of the stream at a time); as many as keep about one CTA per SM busy. 1D outlier streams
  // are one long walk: one CTA each, launched beside the clusters on a second stream.
  int n_cl = 0;
  for (int c = 0; c < nj; c++)
    if (!jobs[c].skip && jobs[c].pow2 && jobs[c].kind != 1)
      n_cl++;
  int n_single = 0;
  for (int c = 0; c < nj; c++)
    if (!jobs[c].skip && jobs[c].pow2 && jobs[c].kind == 1)
      n_single++;
  // every CTA needs an SM of its own (212 KB of shared memory): as large a cluster as still leaves
  // all streams of the batch resident at once (measured: 64 + 64 streams, R = 2 -> the 1D streams
  // queue behind the clusters and the stage takes 67 ms instead of 43)
  int R = 1;
  while (R < kFMaxR && n_cl > 0 && n_cl * R * 2 + n_single <= 148)
    R *= 2;
  if (const char* e = std::getenv("SPERR_B200_DEC_CLUSTER"))
    R = std::max(1, std::min(kFMaxR, std::atoi(e)));
  if (R & (R - 1))
    R = 1;
  bool any_cl = false, any_single = false;
  if (R > 1) {
    w.boxes.reserve(sizeof(ClusterBox) * size_t(n_cl) + 16);
    w.scr.reserve(size_t(n_cl) * R * 2 * kFScr * sizeof(node_t) + 16);
  }
  for (int c = 0, k = 0; c < nj; c++) {
    DecChunk& d = w.h[c];
    d.R = 1;
    if (jobs[c].skip || !jobs[c].pow2)
      continue;
    if (R > 1 && jobs[c].kind != 1) {
      d.R = R;
      d.box = w.boxes.as<ClusterBox>() + k;
      d.scr = w.scr.as<node_t>() + size_t(k) * R * 2 * kFScr;
      k++;
      any_cl = true;
    }
    else
      any_single = true;
  }
  w.dchunks.reserve(sizeof(DecChunk) * nj);
  rt::h2d(w.dchunks.p, w.h.data(), sizeof(DecChunk) * nj, st);
  const size_t aux_bytes = size_t(nj) * 24;
  w.aux.reserve(aux_bytes);
  unsigned char* aux = w.aux.as<unsigned char>();
  rt::h2d(aux, srcs.data(), nj * 8, st);
  rt::h2d(aux + nj * 8, lens.data(), nj * 8, st);
  rt::h2d(aux + nj * 16, words.data(), nj * 8, st);
  DecChunk* dch = w.dchunks.as<DecChunk>();
  {
    rt::ProfScope ps("dec.stage_bits", st);
    const unsigned gx = unsigned(std::min<size_t>((max_words + 255) / 256, 256));
    LAUNCH(k_stage_bits, dim3(gx, nj), dim3(256), 0, st, dch,
           reinterpret_cast<const unsigned char* const*>(aux),
           reinterpret_cast<const unsigned long long*>(aux + nj * 8),
           reinterpret_cast<const unsigned long long*>(aux + nj * 16));
  }
  {
    rt::ProfScope ps("dec.speck_decode", st);
    if (any_slow3)
      LAUNCH(k_speck_decode<DecTree3D>, dim3(nj), dim3(kDecThreads), 0, st, dch, DecTree3D::Data{d_shapes});
    if (any_slow2)
      LAUNCH(k_speck_decode<DecTree2D>, dim3(nj), dim3(kDecThreads), 0, st, dch, DecTree2D::Data{d_shapes});
    if (any_slow1)
      LAUNCH(k_speck_decode<DecTree1D>, dim3(nj), dim3(kDecThreads), 0, st, dch, DecTree1D::Data{0});
    if (any_fast) {
#ifndef SPERR_EMUL
      static rt::OncePerDevice once;
      static cudaStream_t sides[rt::kMaxDevices] = {};
      static cudaEvent_t evs_fork[rt::kMaxDevices] = {}, evs_join[rt::kMaxDevices] = {};
      cudaStream_t& side = sides[rt::cur_dev()];
      cudaEvent_t &ev_fork = evs_fork[rt::cur_dev()], &ev_join = evs_join[rt::cur_dev()];
      if (once.first()) {
        RT_CHECK(cudaFuncSetAttribute(k_speck_decode_fast<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(sizeof(FastSmem))));
        RT_CHECK(cudaFuncSetAttribute(k_speck_decode_fast<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(sizeof(FastSmem))));
        RT_CHECK(cudaStreamCreateWithFlags(&side, cudaStreamNonBlocking));
        RT_CHECK(cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming));
        RT_CHECK(cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming));
      }
      const bool fork = any_cl && any_single;
      if (fork) {   // the single-CTA streams run beside the clusters
        RT_CHECK(cudaEventRecord(ev_fork, st));
        RT_CHECK(cudaStreamWaitEvent(side, ev_fork, 0));
      }
      if (any_cl)
        LAUNCH_CLUSTER(k_speck_decode_fast<true>, dim3(unsigned(nj) * unsigned(R)), dim3(kDecThreads),
                       sizeof(FastSmem), st, unsigned(R), dch, R);
      if (any_single)
        LAUNCH(k_speck_decode_fast<false>, dim3(nj), dim3(kDecThreads), sizeof(FastSmem), fork ? side : st, dch, 1);
      if (fork) {
        RT_CHECK(cudaEventRecord(ev_join, side));
        RT_CHECK(cudaStreamWaitEvent(st, ev_join, 0));
      }
#else
      if (any_cl)
        LAUNCH_CLUSTER(k_speck_decode_fast<true>, dim3(unsigned(nj) * unsigned(R)), dim3(kDecThreads),
                       sizeof(FastSmem), st, unsigned(R), dch, R);
      if (any_single)
        LAUNCH(k_speck_decode_fast<false>, dim3(nj), dim3(kDecThreads), sizeof(FastSmem), st, dch, 1);
#endif
    }
  }
  rt::d2h(w.h.data(), w.dchunks.p, sizeof(DecChunk) * nj, st);
  rt::sync(st);
  if (std::getenv("SPERR_B200_DECPROF"))
    for (int c = 0; c < nj && c < 2; c++)
      if (w.h[c].pow2 && !w.h[c].skip)
        std::fprintf(stderr,
                     "decprof job %d n=%llu: lip %.2f  windows %.2f (%llu)  chains %.2f  expand %.2f  walker "
                     "%.2f  Mcycles\n",
                     c, w.h[c].n, w.h[c].prof[0] * 1e-6, w.h[c].prof[1] * 1e-6, w.h[c].prof[6],
                     w.h[c].prof[2] * 1e-6, w.h[c].prof[3] * 1e-6, w.h[c].prof[4] * 1e-6);
  if (std::getenv("SPERR_B200_DECTRACE"))
    for (int c = 0; c < nj && c < 2; c++)
      if (w.h[c].pow2 && !w.h[c].skip)
        for (int n = w.h[c].planes - 1; n >= 0 && n < kMaxPlanes; n--)
          std::fprintf(stderr, "dectrace job %d plane %d: lip %llu %llx | chains %llu %llx | walk %llu %llx\n", c, n,
                       w.h[c].dbg[n][0][0], w.h[c].dbg[n][0][1], w.h[c].dbg[n][1][0], w.h[c].dbg[n][1][1],
                       w.h[c].dbg[n][2][0], w.h[c].dbg[n][2][1]);
  if (std::getenv("SPERR_B200_DECTRACE"))
    for (int c = 0; c < nj && c < 2; c++)
      if (w.h[c].pow2 && !w.h[c].skip) {
        for (int pi = 0; pi < 2; pi++)
          for (int sgi = 0; sgi < 3; sgi++) {
            std::fprintf(stderr, "dectrace job %d plane#%d stage %d lists:", c, pi, sgi);
            for (int l = 0; l < w.h[c].nlis && l < 32; l++)
              std::fprintf(stderr, " %u", w.h[c].dbgl[pi][sgi][l]);
            std::fprintf(stderr, "\n");
          }
        for (unsigned k = 0; k < w.h[c].dbgw_n && k < 16; k++)
          std::fprintf(stderr, "dectrace job %d walk: plane %u depth %u roots %u visited %u survivors %u lis %u q %u staged %u\n",
                       c, w.h[c].dbgw[k][0], w.h[c].dbgw[k][1], w.h[c].dbgw[k][2], w.h[c].dbgw[k][3],
                       w.h[c].dbgw[k][4], w.h[c].dbgw[k][5], w.h[c].dbgw[k][6], w.h[c].dbgw[k][7]);
      }
  for (int c = 0; c < nj; c++)
    if (w.h[c].err)
      throw std::runtime_error("SPECK decoder: list capacity exceeded (corrupt stream?)");
}

}  // namespace sperr_b200
