// Chunk-batch decompression: the GPU counterpart of SPERR3D_OMP_D's per-chunk loop
// (/root/reference/src/SPERR3D_OMP_D.cpp:51-135) with SPECK_FLT::use_bitstream / decompress
// (src/SPECK_FLT.cpp:27-109, 543-606) inside.
#include "pipeline.h"

namespace sperr_b200 {

namespace {

// Device footprint per value while decoding: fp64 buffer 8, magnitudes 4 (+4 for outliers),
// masks ~1, lists ~1.2.
size_t pick_dec_batch(const std::vector<Chunk>& chunks, size_t first)
{
  size_t free_b = size_t(64) << 30, total_b = 0;
#ifndef SPERR_EMUL
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
    free_b = size_t(64) << 30;
  free_b += rt::DBuf::held();   // our own grow-only buffers are reused, not allocated again
#endif
  (void)total_b;
  const double budget = double(free_b) * 0.7;
  double used = 0;
  size_t n = 0;
  while (first + n < chunks.size() && n < 1024) {
    const double need = double(chunks[first + n].nelem()) * 24.0 + (16 << 20);
    if (n > 0 && used + need > budget)
      break;
    used += need;
    n++;
  }
  return std::max<size_t>(n, 1);
}

}  // namespace

// out[i * stride .. + n[i]) = src[off[i] .. + n[i]): the few header bytes of every chunk stream
__global__ void k_gather_bytes(const uint8_t* src, const unsigned long long* off, const unsigned* n, int stride,
                               uint8_t* out)
{
  const unsigned i = blockIdx.x;
  for (unsigned b = threadIdx.x; b < n[i]; b += blockDim.x)
    out[(size_t)i * stride + b] = src[off[i] + b];
}

void Decompressor::decompress(const uint8_t* h_stream, const uint8_t* d_stream,
                              const std::vector<Chunk>& chunks, const std::vector<ChunkStream>& cs,
                              const SrcVol& dst, cudaStream_t st, bool is_2d)
{
  size_t first = 0;
  while (first < chunks.size()) {
    size_t nb = pick_dec_batch(chunks, first);
    if (max_batch)
      nb = std::min(nb, max_batch);
    std::vector<Chunk> sub(chunks.begin() + first, chunks.begin() + first + nb);
    batch_first_ = first;
    whole_call_ = nb == chunks.size();   // group reporting needs slab-aligned groups: one batch only
    run_batch(h_stream, d_stream, sub, cs.data() + first, dst, st, is_2d);
    if (after_batch)
      after_batch(first, nb);
    first += nb;
  }
}

void Decompressor::run_batch(const uint8_t* h_stream, const uint8_t* d_stream,
                             const std::vector<Chunk>& chunks, const ChunkStream* cs,
                             const SrcVol& dst, cudaStream_t st, bool is_2d)
{
  const int nc = int(chunks.size());

  // ---- parse the per-chunk headers (SPECK_FLT::use_bitstream) ----
  struct Parsed {
    bool is_const = false;
    double mean = 0, q = 0, cval = 0;
    int planes = 0;
    unsigned long long total_bits = 0;
    size_t spk_off = 0, spk_bytes = 0;     // payload (after the 9-byte header), container offsets
    bool has_out = false;
    int oplanes = 0;
    unsigned long long ototal = 0;
    size_t out_off = 0, out_bytes = 0;
  };
  std::vector<Parsed> ps(nc);
  bool any_wide = false, any_out = false, any_owide = false;
  // Without a host copy of the streams the bytes the parser reads are fetched from the device: the
  // first 26 bytes of every chunk (conditioner + SPECK header), then the 9-byte header of the
  // outlier stream where one may follow.
  std::vector<uint8_t> hdr1, hdr2;
  if (!h_stream) {
    hdr1.assign(size_t(nc) * 26, 0);
    hdr2.assign(size_t(nc) * 9, 0);
    std::vector<unsigned long long> off(nc);
    std::vector<unsigned> cnt(nc);
    hdr_off_.reserve(size_t(nc) * 8);
    hdr_cnt_.reserve(size_t(nc) * 4);
    hdr_out_.reserve(size_t(nc) * 26);
    auto fetch = [&](int stride, uint8_t* dst) {
      rt::h2d(hdr_off_.p, off.data(), size_t(nc) * 8, st);
      rt::h2d(hdr_cnt_.p, cnt.data(), size_t(nc) * 4, st);
      LAUNCH(k_gather_bytes, dim3(unsigned(nc)), dim3(32), 0, st, d_stream, hdr_off_.as<unsigned long long>(),
             hdr_cnt_.as<unsigned>(), stride, hdr_out_.as<uint8_t>());
      rt::d2h(dst, hdr_out_.p, size_t(nc) * stride, st);
      rt::sync(st);
    };
    for (int c = 0; c < nc; c++) {
      off[c] = cs[c].off;
      cnt[c] = unsigned(std::min<size_t>(26, cs[c].len));
    }
    fetch(26, hdr1.data());
    bool any2 = false;
    for (int c = 0; c < nc; c++) {
      const uint8_t* p = &hdr1[size_t(c) * 26];
      cnt[c] = 0;
      off[c] = cs[c].off;
      if (cs[c].len < 26 || (p[0] & 0x01))
        continue;
      unsigned long long tb;
      std::memcpy(&tb, p + 18, 8);
      if (tb > (1ull << 48))
        continue;
      const size_t pos = 17 + std::min(9 + size_t((tb + 7) / 8), cs[c].len - 17);
      if (pos < cs[c].len && cs[c].len - pos >= 9) {
        off[c] = cs[c].off + pos;
        cnt[c] = 9;
        any2 = true;
      }
    }
    if (any2)
      fetch(9, hdr2.data());
  }
  for (int c = 0; c < nc; c++) {
    const uint8_t* p = h_stream ? h_stream + cs[c].off : &hdr1[size_t(c) * 26];
    const size_t len = cs[c].len;
    Parsed& P = ps[c];
    if (len < 17)
      throw std::runtime_error("chunk stream shorter than the conditioner header");
    if (p[0] & 0x01) {  // constant field: Conditioner::inverse_condition, src/Conditioner.cpp:66-80
      if (len != 17)
        throw std::runtime_error("constant chunk with trailing bytes");
      P.is_const = true;
      std::memcpy(&P.cval, p + 9, 8);
      continue;
    }
    std::memcpy(&P.mean, p + 1, 8);
    std::memcpy(&P.q, p + 9, 8);
    size_t remaining = len - 17;
    if (remaining < 9)
      throw std::runtime_error("chunk stream without a SPECK header");
    const uint8_t* sp = p + 17;
    P.planes = sp[0];
    std::memcpy(&P.total_bits, sp + 1, 8);
    if (P.total_bits > (1ull << 48))
      throw std::runtime_error("implausible SPECK stream length");
    const size_t full = 9 + size_t((P.total_bits + 7) / 8);
    const size_t speck_len = std::min(full, remaining);
    P.spk_off = cs[c].off + 17 + 9;
    P.spk_bytes = speck_len - 9;
    any_wide |= P.planes > 32;
    size_t pos = 17 + speck_len;
    if (pos < len) {
      remaining = len - pos;
      if (remaining >= 9) {
        const uint8_t* op = h_stream ? p + pos : &hdr2[size_t(c) * 9];
        unsigned long long nb;
        std::memcpy(&nb, op + 1, 8);
        const size_t ofull = nb > (1ull << 48) ? 0 : 9 + size_t((nb + 7) / 8);
        if (remaining == ofull) {
          P.has_out = true;
          P.oplanes = op[0];
          P.ototal = nb;
          P.out_off = cs[c].off + pos + 9;
          P.out_bytes = ofull - 9;
          any_out = true;
          any_owide |= P.oplanes > 32;
        }
      }
    }
  }

  (void)any_wide;
  (void)any_owide;
  b_.setup(chunks, true, false, false, st, false, is_2d);
  for (int c = 0; c < nc; c++) {
    ChunkDev& d = b_.h[c];
    d.is_const = ps[c].is_const ? 1 : 0;
    d.first_val = ps[c].cval;
    d.mean = ps[c].mean;
    d.q = ps[c].q;
  }
  b_.push(st);

  // ---- SPECK3D decode: sorting passes per chunk, then magnitudes + de-quantisation grid-wide ----
  // jobs [0, nc): the chunks' SPECK3D streams; jobs [nc, 2 nc): their SPECK1D outlier streams. All
  // of them are decoded side by side (one CTA each), so the outlier streams cost no extra time.
  std::vector<DecJob> jobs(2 * nc);
  for (int c = 0; c < nc; c++) {
    DecJob& j = jobs[c];
    const ShapeTables& sh = b_.shapes[b_.h[c].shape];
    j.skip = ps[c].is_const || ps[c].planes == 0;
    j.n = b_.h[c].n;
    j.shape = b_.h[c].shape;
    j.kind = is_2d ? 2 : 0;
    j.d_payload = d_stream + ps[c].spk_off;
    j.payload_bytes = ps[c].spk_bytes;
    j.planes = ps[c].planes;
    j.total_bits = ps[c].total_bits;
    j.nlis = sh.h.nlis;
    // lis_off lives inside the ShapeHeader copy on the device
    j.d_lis_off = nullptr;
    j.lis_total = sh.h.lis_off[kMaxLis];
    const int J3 = std::max(sh.h.ax[0].D, std::max(sh.h.ax[1].D, sh.h.ax[2].D));
    auto is_pow2 = [](uint32_t v) { return v != 0 && (v & (v - 1)) == 0; };
    if (is_2d && is_pow2(sh.h.nx) && is_pow2(sh.h.ny) && J3 >= 4 && sh.h.nxf2d >= 1 &&
        !std::getenv("SPERR_B200_NO_FASTDEC")) {
      // power-of-two slice: quadtree of aligned boxes + the set I, decoded by k_speck_decode_fast
      j.pow2 = 1;
      j.Dx = sh.h.ax[0].D; j.Dy = sh.h.ax[1].D; j.Dz = 0;
      j.nx = sh.h.nx; j.ny = sh.h.ny;
      j.nroots = 1;
      j.roots[0] = (unsigned long long)sh.h.nxf2d << 32;   // the approximation band, index 0
      j.nxf2d = sh.h.nxf2d;
    }
    else if (sh.h.pow2 && J3 >= 4 && !std::getenv("SPERR_B200_NO_FASTDEC")) {
      // power-of-two extents: every set is an aligned box, decoded by k_speck_decode_fast
      j.pow2 = 1;
      j.Dx = sh.h.ax[0].D; j.Dy = sh.h.ax[1].D; j.Dz = sh.h.ax[2].D;
      j.nx = sh.h.nx; j.ny = sh.h.ny;
      j.nroots = sh.h.nroots;
      for (int r = 0; r < sh.h.nroots; r++) {
        const RootDesc& rd = sh.h.roots[r];
        const int dj = sh.h.lv[rd.level].j;   // single chain: depth of the set
        const int bx = std::min(dj, j.Dx), by = std::min(dj, j.Dy);
        j.roots[r] = ((unsigned long long)dj << 32) |
                     (unsigned long long)(unsigned(rd.ix) | (unsigned(rd.iy) << bx) | (unsigned(rd.iz) << (bx + by)));
      }
    }
  }
  {
    // device addresses of ShapeHeader::lis_off: fetch the ShapeDev array once
    std::vector<ShapeDev> sd(b_.shapes.size());
    rt::d2h(sd.data(), b_.d_shapes.p, sizeof(ShapeDev) * sd.size(), st);
    rt::sync(st);
    for (int c = 0; c < nc; c++)
      jobs[c].d_lis_off = reinterpret_cast<const unsigned long long*>(
          reinterpret_cast<const unsigned char*>(sd[b_.h[c].shape].h) + offsetof(ShapeHeader, lis_off));
  }
  // ---- outlier streams (SPECK1D, src/SPECK_FLT.cpp:576-585) ----
  std::vector<double> tols(nc, 0.0);
  std::vector<unsigned long long> h_lis_off;
  if (any_out) {
    std::vector<size_t> lo(nc, 0);
    std::vector<int> nl(nc, 0);
    for (int c = 0; c < nc; c++) {
      if (!ps[c].has_out || ps[c].oplanes == 0)
        continue;
      // list capacities: level l holds at most min(2^l, bits in the stream) live sets
      const int nlis = int(num_of_partitions(b_.h[c].n)) + 2;
      nl[c] = nlis;
      lo[c] = h_lis_off.size();
      unsigned long long acc = 0;
      for (int l = 0; l <= nlis; l++) {
        h_lis_off.push_back(acc);
        const unsigned long long cap2 = l >= 40 ? ~0ull : (1ull << l);
        acc += std::min<unsigned long long>(cap2, ps[c].ototal + 2);
      }
    }
    lis_off1_.reserve(h_lis_off.size() * 8 + 16);
    rt::h2d(lis_off1_.p, h_lis_off.data(), h_lis_off.size() * 8, st);
    for (int c = 0; c < nc; c++) {
      DecJob& j = jobs[nc + c];
      j.kind = 1;
      j.skip = !ps[c].has_out || ps[c].oplanes == 0;
      if (j.skip)
        continue;
      j.n = b_.h[c].n;
      j.d_payload = d_stream + ps[c].out_off;
      j.payload_bytes = ps[c].out_bytes;
      j.planes = ps[c].oplanes;
      j.total_bits = ps[c].ototal;
      j.nlis = nl[c];
      j.d_lis_off = lis_off1_.as<unsigned long long>() + lo[c];
      j.lis_total = h_lis_off[lo[c] + nl[c]];
      const unsigned long long nn = b_.h[c].n;
      if (nn >= 16 && (nn & (nn - 1)) == 0 && !std::getenv("SPERR_B200_NO_FASTDEC")) {
        // power-of-two length: the binary tree of SPECK1D is the fast decoder's tree with one axis;
        // the two initial halves sit at depth 1 (src/SPECK1D_INT.cpp:18-56)
        j.pow2 = 1;
        j.Dx = 0;
        while ((1ull << j.Dx) < nn)
          j.Dx++;
        j.Dy = j.Dz = 0;
        j.nx = unsigned(nn);
        j.ny = 1;
        j.nroots = 2;
        j.roots[0] = (1ull << 32) | 0ull;
        j.roots[1] = (1ull << 32) | 1ull;
      }
      tols[c] = ps[c].q / 1.5;  // src/SPECK_FLT.cpp:578
    }
    tols_.reserve(nc * 8);
    rt::h2d(tols_.p, tols.data(), nc * 8, st);
  }
  {
    rt::ProfScope pd("d.speck", st);
    speck_decode(w_, jobs, b_.dev_shapes(), st);
  }
  {
    rt::ProfScope pq("d.reconstruct", st);
    w_.fill_n = b_.max_n;
    speck_reconstruct(w_, b_.dev(), 0, nullptr, 0, nc, OutlierSink{}, st);
  }

  // ---- outlier correctors: sorted (chunk, position) list + one flag bit per value ----
  CorrectorList cor{nullptr, nullptr, nullptr};
  unsigned long long ncor = 0;
  if (any_out) {
    rt::ProfScope po("d.outliers", st);
    size_t words = 0, cap = 16;
    for (int c = 0; c < nc; c++) {
      words += (b_.h[c].n + 31) / 32 + 1;
      cap += size_t(ps[c].ototal / 2 + 1);   // a corrector costs at least a significance + a sign bit
    }
    obits_.reserve(words * 4);
    rt::dset(obits_.p, 0, words * 4, st);
    size_t wo = 0;
    for (int c = 0; c < nc; c++) {
      b_.h[c].obits = obits_.as<uint32_t>() + wo;
      wo += (b_.h[c].n + 31) / 32 + 1;
    }
    b_.push(st);
    ckey_[0].reserve(cap * 8);
    ckey_[1].reserve(cap * 8);
    cval_[0].reserve(cap * 8);
    cval_[1].reserve(cap * 8);
    ccount_.reserve(8 + size_t(nc + 1) * 8);
    rt::dset(ccount_.p, 0, 8 + size_t(nc + 1) * 8, st);
    OutlierSink sink;
    sink.total = ccount_.as<unsigned long long>();
    sink.per_chunk = reinterpret_cast<unsigned*>(ccount_.as<unsigned char>() + 8);
    sink.key = ckey_[0].as<unsigned long long>();
    sink.err = cval_[0].as<double>();
    sink.cap = cap;
    w_.fill_n = 0;
    speck_reconstruct(w_, b_.dev(), 1, tols_.as<double>(), nc, nc, sink, st);
    std::vector<unsigned char> hc(8 + size_t(nc) * 4);
    rt::d2h(hc.data(), ccount_.p, hc.size(), st);
    rt::sync(st);
    std::memcpy(&ncor, hc.data(), 8);
    if (ncor > cap)
      throw std::runtime_error("outlier corrector list overflow");
    if (ncor) {
      std::vector<unsigned long long> off(nc + 1, 0);
      for (int c = 0; c < nc; c++) {
        unsigned n;
        std::memcpy(&n, hc.data() + 8 + size_t(c) * 4, 4);
        off[c + 1] = off[c] + n;
      }
      coff_.reserve((nc + 1) * 8);
      rt::h2d(coff_.p, off.data(), (nc + 1) * 8, st);
      const size_t tb = sort_tmp_bytes(size_t(ncor));
      csort_.reserve(tb);
      int bits = 33;
      while ((1ll << (bits - 32)) < nc)
        bits++;
      sort_pairs_u64(ckey_[0].as<unsigned long long>(), ckey_[1].as<unsigned long long>(),
                     cval_[0].as<unsigned long long>(), cval_[1].as<unsigned long long>(), size_t(ncor), bits,
                     csort_.p, tb, st);
      rt::sync(st);   // `off` leaves scope
      cor.key = ckey_[1].as<unsigned long long>();
      cor.val = cval_[1].as<double>();
      cor.off = coff_.as<unsigned long long>();
    }
  }

  // ---- inverse transform ----
  std::vector<std::vector<int>> groups(b_.shapes.size());
  bool any_unfused = false;
  for (int c = 0; c < nc; c++) {
    if (!ps[c].is_const)
      groups[b_.h[c].shape].push_back(c);
    if (ps[c].is_const || !b_.h[c].fused)
      any_unfused = true;
  }
  std::vector<int> flat;
  std::vector<size_t> goff;
  for (auto& g : groups) {
    goff.push_back(flat.size());
    flat.insert(flat.end(), g.begin(), g.end());
  }
  ids_.reserve(flat.size() * 4 + 4);
  rt::h2d(ids_.p, flat.data(), flat.size() * 4, st);
  rt::sync(st);
  // The inverse transform writes final values. When every chunk goes through the fused kernels
  // (which scatter into `dst` themselves) it is run slab group by slab group, and the caller is told
  // after each group (with an event) so that it can start moving that part of the result while the
  // next group is transformed.
  groups_posted = false;
  const bool by_groups =
      after_group && group_chunks > 0 && !any_unfused && size_t(nc) > group_chunks && whole_call_;
  {
    rt::ProfScope pt("d.idwt", st);
    for (size_t g0 = 0; g0 < size_t(nc); g0 += by_groups ? group_chunks : size_t(nc)) {
      const size_t g1 = by_groups ? std::min(size_t(nc), g0 + group_chunks) : size_t(nc);
      for (size_t s = 0; s < groups.size(); s++) {
        if (groups[s].empty())
          continue;
        // chunk ids of a shape group ascend: the members inside [g0, g1) are one sub-range
        const auto lo = std::lower_bound(groups[s].begin(), groups[s].end(), int(g0));
        const auto hi = std::lower_bound(groups[s].begin(), groups[s].end(), int(g1));
        if (lo == hi)
          continue;
        const ShapeHeader& h = b_.shapes[s].h;
        const int* ids = ids_.as<int>() + goff[s] + (lo - groups[s].begin());
        const int cnt = int(hi - lo);
        if (b_.h[groups[s][0]].fused)   // corrector, mean, conversion and scatter fused into level 0
          launch_dwt_fused_inverse(dst, 1, b_.dev(), ids, cnt, h.nx, h.ny, h.nz, 0.0, OutlierSink{}, cor, st);
        else
          launch_dwt(true, b_.dev(), ids, cnt, h.nx, h.ny, h.nz, is_2d, st);
      }
#ifndef SPERR_EMUL
      if (by_groups) {
        cudaEvent_t ev;
        RT_CHECK(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        RT_CHECK(cudaEventRecord(ev, st));
        after_group(batch_first_ + g0, g1 - g0, ev);   // takes ownership of the event
        groups_posted = true;
      }
#endif
    }
  }

  // ---- multi-resolution output: the coarse boxes the fused inverse transform left behind ----
  if (multires) {
    if (any_unfused && nc > 0 && !b_.h[0].fused)
      throw std::runtime_error("multi-resolution decoding needs the fused transform path");
    const ShapeHeader& h0 = b_.shapes[b_.h[0].shape].h;
    for (size_t h = 0; h < multires->d_level.size(); h++)
      launch_level_gather(b_.dev(), nc, h0.nx, h0.ny, h0.nz, int(h), multires->d_level[h], multires->is_float,
                          multires->dims[h][0], multires->dims[h][1], st);
  }

  // ---- the other chunks: correctors, mean, conversion, scatter; constant chunks: fill ----
  if (any_unfused) {
    rt::ProfScope psc("d.scatter", st);
    if (ncor)
      launch_apply_correctors(b_.dev(), cor.key, cor.val, ncor, st);
    launch_scatter_out(dst, b_.dev(), nc, b_.max_n, st);
  }
  rt::sync(st);
}

}  // namespace sperr_b200
