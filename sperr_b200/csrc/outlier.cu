// PWE-mode outlier path: detection (compaction in raster order) and the SPECK1D outlier coder.
//
// Reproduces, bit for bit:
//   outlier scan                    /root/reference/src/SPECK_FLT.cpp:468-474
//   Outlier_Coder::encode           /root/reference/src/Outlier_Coder.cpp:71-131 (+ m_quantize :188-204)
//   SPECK1D_INT / _ENC              /root/reference/src/SPECK1D_INT.cpp:18-56, src/SPECK1D_INT_ENC.cpp:12-177
// The coded array is as long as the chunk but holds only a few thousand non-zeros, so instead of
// a dense array the binary partition tree is materialised only along the paths that lead to an
// outlier (plus the empty siblings that the coder still has to test every plane); the
// tree-agnostic engine of speck_engine.cuh then does the rest.
#include "outlier.h"

#include "speck_engine.cuh"

namespace sperr_b200 {

// ---------------------------------------------------------------------------------------------
// detection
// ---------------------------------------------------------------------------------------------

constexpr int kOutBlock = 1024;

__device__ __forceinline__ double out_diff(const SrcVol& src, const ChunkDev& ch,
                                           unsigned long long e)
{
  const unsigned x = unsigned(e % ch.nx);
  const unsigned y = unsigned((e / ch.nx) % ch.ny);
  const unsigned z = unsigned(e / ((unsigned long long)ch.nx * ch.ny));
  const unsigned long long si = (unsigned long long)(ch.z0 + z) * src.vx * src.vy +
                                (unsigned long long)(ch.y0 + y) * src.vx + (ch.x0 + x);
  const double v = src.is_float ? double(reinterpret_cast<const float*>(src.ptr)[si])
                                : reinterpret_cast<const double*>(src.ptr)[si];
  const double orig = __dsub_rn(v, ch.mean);
  return __dsub_rn(orig, ch.coef[e]);
}

// counts[c * nblk + blk] = number of outliers in that block of 1024 values
__global__ void k_outlier_count(SrcVol src, const ChunkDev* chunks, double tol, unsigned* counts,
                                unsigned nblk)
{
  const unsigned c = blockIdx.y, blk = blockIdx.x;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || (unsigned long long)blk * kOutBlock >= ch.n)
    return;
  __shared__ unsigned s_cnt;
  if (threadIdx.x == 0)
    s_cnt = 0;
  __syncthreads();
  const unsigned long long e = (unsigned long long)blk * kOutBlock + threadIdx.x;
  bool is_out = false;
  if (e < ch.n)
    is_out = fabs(out_diff(src, ch, e)) > tol;
  const unsigned b = __ballot_sync(0xffffffffu, is_out);
  if ((threadIdx.x & 31) == 0 && b)
    atomicAdd(&s_cnt, unsigned(__popc(b)));
  __syncthreads();
  if (threadIdx.x == 0)
    counts[(size_t)c * nblk + blk] = s_cnt;
}

// offs = exclusive scan of counts over the whole (chunk-major) array
__global__ void k_outlier_write(SrcVol src, const ChunkDev* chunks, double tol,
                                const unsigned long long* offs, unsigned nblk, unsigned* opos,
                                double* oerr)
{
  const unsigned c = blockIdx.y, blk = blockIdx.x;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || (unsigned long long)blk * kOutBlock >= ch.n)
    return;
  __shared__ unsigned s_w[32];
  const unsigned long long e = (unsigned long long)blk * kOutBlock + threadIdx.x;
  double d = 0.0;
  bool is_out = false;
  if (e < ch.n) {
    d = out_diff(src, ch, e);
    is_out = fabs(d) > tol;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b = __ballot_sync(0xffffffffu, is_out);
  if (lane == 0)
    s_w[warp] = __popc(b);
  __syncthreads();
  if (warp == 0) {
    unsigned v = s_w[lane], inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o)
        inc += t;
    }
    s_w[lane] = inc - v;
  }
  __syncthreads();
  if (is_out) {
    const unsigned long long o = offs[(size_t)c * nblk + blk] + s_w[warp] + __popc(b & ((1u << lane) - 1));
    opos[o] = unsigned(e);
    oerr[o] = d;
  }
}

__global__ void k_pick_offsets(const unsigned long long* offs, unsigned nblk, int nchunks,
                               unsigned long long* out)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c <= nchunks)
    out[c] = offs[(size_t)c * nblk];
}

void OutlierCoder::detect(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                          double tol, cudaStream_t st)
{
  const unsigned nblk = unsigned((max_n + kOutBlock - 1) / kOutBlock);
  const size_t ncnt = (size_t)nchunks * nblk;
  cnt_.reserve(ncnt * 4);
  offs_.reserve((ncnt + 2) * 8);
  scan_tmp_.reserve(scan_tmp_bytes(ncnt + 1));
  rt::dset(cnt_.p, 0, ncnt * 4, st);
  unsigned* d_cnt = cnt_.as<unsigned>();
  unsigned long long* d_offs = offs_.as<unsigned long long>();
  LAUNCH(k_outlier_count, dim3(nblk, nchunks), dim3(kOutBlock), 0, st, src, d_chunks, tol, d_cnt, nblk);
  exclusive_scan_u32(d_cnt, d_offs, ncnt, scan_tmp_.p, st);
  // per-chunk offsets are every nblk-th entry of the scan
  pick_.reserve((nchunks + 1) * 8);
  unsigned long long* d_pick = pick_.as<unsigned long long>();
  LAUNCH(k_pick_offsets, dim3((nchunks + 64) / 64), dim3(64), 0, st, d_offs, nblk, nchunks, d_pick);
  ooff.assign(nchunks + 1, 0);
  rt::d2h(ooff.data(), d_pick, (nchunks + 1) * 8, st);
  rt::sync(st);
  const size_t total = size_t(ooff[nchunks]);
  opos_.reserve((total + 1) * 4);
  oerr_.reserve((total + 1) * 8);
  if (total) {
    unsigned* d_pos = opos_.as<unsigned>();
    double* d_err = oerr_.as<double>();
    LAUNCH(k_outlier_write, dim3(nblk, nchunks), dim3(kOutBlock), 0, st, src, d_chunks, tol, d_offs, nblk,
           d_pos, d_err);
  }
}

// ---- detection through an unordered append list (fused inverse transform + k_outlier_append) ----

// Outliers of the chunks that are NOT handled by the fused inverse transform (their reconstruction
// sits in coef after the in-place inverse transform).
__global__ void k_outlier_append(SrcVol src, const ChunkDev* chunks, double tol, OutlierSink sink)
{
  const unsigned c = blockIdx.y;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || ch.fused)
    return;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ch.n;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const double d = out_diff(src, ch, e);
    if (fabs(d) > tol)
      outlier_append(sink, c, e, d);
  }
}

__global__ void k_outlier_split(const unsigned long long* key, unsigned* opos, unsigned long long n)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n)
    opos[i] = unsigned(key[i] & 0xffffffffull);
}

OutlierSink OutlierCoder::begin_detect(int nchunks, size_t total_values, cudaStream_t st)
{
  if (sink_cap_ == 0)
    sink_cap_ = std::max<size_t>(size_t(1) << 16, total_values / 64);
  skey_[0].reserve(sink_cap_ * 8);
  serr_[0].reserve(sink_cap_ * 8);
  scount_.reserve(8 + size_t(nchunks) * 4);
  rt::dset(scount_.p, 0, 8 + size_t(nchunks) * 4, st);
  OutlierSink s;
  s.total = scount_.as<unsigned long long>();
  s.per_chunk = reinterpret_cast<unsigned*>(scount_.as<unsigned char>() + 8);
  s.key = skey_[0].as<unsigned long long>();
  s.err = serr_[0].as<double>();
  s.cap = sink_cap_;
  return s;
}

void OutlierCoder::append_unfused(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                                  double tol, const OutlierSink& sink, cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 2048));
  LAUNCH(k_outlier_append, dim3(gx, nchunks), dim3(256), 0, st, src, d_chunks, tol, sink);
}

bool OutlierCoder::end_detect(int nchunks, cudaStream_t st)
{
  std::vector<unsigned char> h(8 + size_t(nchunks) * 4);
  rt::d2h(h.data(), scount_.p, h.size(), st);
  rt::sync(st);
  unsigned long long total;
  std::memcpy(&total, h.data(), 8);
  if (total > sink_cap_) {   // the list was too short: the caller repeats the detection
    sink_cap_ = size_t(total) + size_t(total) / 8 + 1024;
    return false;
  }
  ooff.assign(nchunks + 1, 0);
  for (int c = 0; c < nchunks; c++) {
    unsigned n;
    std::memcpy(&n, h.data() + 8 + size_t(c) * 4, 4);
    ooff[c + 1] = ooff[c] + n;
  }
  opos_.reserve((size_t(total) + 1) * 4);
  oerr_.reserve((size_t(total) + 1) * 8);
  if (total == 0)
    return true;
  // raster order inside every chunk, chunks in order: sort by (chunk << 32 | position)
  skey_[1].reserve(sink_cap_ * 8);
  const size_t tb = sort_tmp_bytes(size_t(total));
  sort_tmp_.reserve(tb);
  int bits = 33;
  while ((1ll << (bits - 32)) < nchunks)
    bits++;
  sort_pairs_u64(skey_[0].as<unsigned long long>(), skey_[1].as<unsigned long long>(),
                 serr_[0].as<unsigned long long>(), oerr_.as<unsigned long long>(), size_t(total), bits,
                 sort_tmp_.p, tb, st);
  LAUNCH(k_outlier_split, dim3(unsigned((total + 255) / 256)), dim3(256), 0, st,
         skey_[1].as<unsigned long long>(), opos_.as<unsigned>(), total);
  return true;
}

void OutlierCoder::set_outliers(const std::vector<unsigned long long>& offsets, const unsigned* h_pos,
                                const double* h_err, cudaStream_t st)
{
  ooff = offsets;
  const size_t total = size_t(ooff.back());
  opos_.reserve((total + 1) * 4);
  oerr_.reserve((total + 1) * 8);
  rt::h2d(opos_.p, h_pos, total * 4, st);
  rt::h2d(oerr_.p, h_err, total * 8, st);
  rt::sync(st);
}

// ---------------------------------------------------------------------------------------------
// quantisation of the outliers (Outlier_Coder::encode :84-102, m_quantize :188-204)
// ---------------------------------------------------------------------------------------------

struct OutMeta {             // one per chunk
  unsigned long long off0, off1;   // range in the concatenated outlier arrays
  unsigned long long maxerr_bits;  // bit pattern of max |err|
  unsigned long long total_len;
  int width;                       // 1, 2, 4, 8; 0 = FE_INVALID
  unsigned root;                   // index of the depth-0 node
  unsigned nlis;                   // num_of_partitions(total_len) + 1
  unsigned npix;                   // pixels created by the coder (LIP / refinement domain)
  unsigned long long pix0;         // first entry of the chunk in the compact pixel arrays
};

__global__ void k_out_maxerr(OutMeta* meta, const double* oerr, int nchunks)
{
  const unsigned c = blockIdx.y;
  OutMeta& m = meta[c];
  const unsigned long long n = m.off1 - m.off0;
  unsigned long long best = 0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long b =
        (unsigned long long)__double_as_longlong(oerr[m.off0 + i]) & 0x7fffffffffffffffull;
    best = b > best ? b : best;
  }
  if (best)
    atomicMax(&m.maxerr_bits, best);
  (void)nchunks;
}

__global__ void k_out_width(OutMeta* meta, int nchunks)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks)
    return;
  OutMeta& m = meta[c];
  const double mx = __longlong_as_double((long long)m.maxerr_bits);
  if (!(mx < 9223372036854775807.0)) {
    m.width = 0;
    return;
  }
  const long long mi = __double2ll_rn(mx);  // quirk: not divided by the tolerance (:89)
  m.width = mi <= 0xFF ? 1 : mi <= 0xFFFF ? 2 : mi <= 0xFFFFFFFFll ? 4 : 8;
}

__global__ void k_out_quant(const OutMeta* meta, const double* oerr, double tol,
                            unsigned long long* omag, unsigned char* osign)
{
  const unsigned c = blockIdx.y;
  const OutMeta& m = meta[c];
  const unsigned long long n = m.off1 - m.off0;
  const double inv = __ddiv_rn(1.0, tol);
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const long long ll = __double2ll_rn(__dmul_rn(oerr[m.off0 + i], inv));
    unsigned long long mag = (unsigned long long)(ll < 0 ? -ll : ll);
    if (m.width == 1)
      mag &= 0xFFull;
    else if (m.width == 2)
      mag &= 0xFFFFull;
    else if (m.width == 4)
      mag &= 0xFFFFFFFFull;
    omag[m.off0 + i] = mag;
    osign[m.off0 + i] = ll >= 0 ? 1 : 0;
  }
}

// ---------------------------------------------------------------------------------------------
// sparse partition tree
// ---------------------------------------------------------------------------------------------

struct ONode {
  unsigned start, len;   // interval of the 1D array
  unsigned a, b;         // outliers inside: indices [a, b) of the concatenated arrays
  unsigned fc;           // first child (children are fc, fc + 1); 0 = none
  unsigned d;            // expansion size; for pixels: index in the compact pixel arrays
  short chunk;
  signed char p;         // msb of the largest magnitude inside, -1 if none
  signed char cpl;       // pixels: plane at which the parent splits, -1 never
  unsigned char depth;
  unsigned char pad[3];
};

struct TreeBuild {
  ONode* nodes;
  unsigned long long cap;
  unsigned* lvl_base;    // [64]
  unsigned* lvl_count;   // [64]
  unsigned* err;
  const unsigned* opos;
  const unsigned long long* omag;
};

__global__ void k_tree_roots(TreeBuild t, OutMeta* meta, int nchunks)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c == 0) {
    t.lvl_base[0] = 0;
    t.lvl_count[0] = unsigned(nchunks);
    t.lvl_base[1] = unsigned(nchunks);
  }
  if (c >= nchunks)
    return;
  ONode n;
  n.start = 0;
  n.len = unsigned(meta[c].total_len);
  n.a = unsigned(meta[c].off0);
  n.b = unsigned(meta[c].off1);
  n.fc = 0;
  n.d = 0;
  n.chunk = short(c);
  n.p = -1;
  n.cpl = -1;
  n.depth = 0;
  t.nodes[c] = n;
  meta[c].root = unsigned(c);
}

// Creates the children of every non-empty set of depth `d`.
__global__ void k_tree_split(TreeBuild t, int d)
{
  const unsigned base = t.lvl_base[d], cnt = t.lvl_count[d], cbase = t.lvl_base[d + 1];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    ONode& n = t.nodes[base + i];
    if (n.b == n.a || n.len <= 1)
      continue;
    const unsigned slot = cbase + atomicAdd(&t.lvl_count[d + 1], 2u);
    if ((unsigned long long)slot + 2 > t.cap) {
      atomicOr(t.err, 1u);
      continue;
    }
    const unsigned len0 = n.len - n.len / 2;
    const unsigned split = n.start + len0;
    unsigned lo = n.a, hi = n.b;  // first outlier with pos >= split
    while (lo < hi) {
      const unsigned mid = (lo + hi) >> 1;
      if (t.opos[mid] < split)
        lo = mid + 1;
      else
        hi = mid;
    }
    ONode c0, c1;
    c0.start = n.start; c0.len = len0; c0.a = n.a; c0.b = lo;
    c1.start = split; c1.len = n.len / 2; c1.a = lo; c1.b = n.b;
    c0.fc = c1.fc = 0;
    c0.d = c1.d = 0;
    c0.chunk = c1.chunk = n.chunk;
    c0.p = c1.p = -1;
    c0.cpl = c1.cpl = -1;
    c0.depth = c1.depth = (unsigned char)(d + 1);
    t.nodes[slot] = c0;
    t.nodes[slot + 1] = c1;
    n.fc = slot;
  }
}

__global__ void k_tree_next(TreeBuild t, int d)
{
  if (threadIdx.x == 0 && blockIdx.x == 0)
    t.lvl_base[d + 2] = t.lvl_base[d + 1] + t.lvl_count[d + 1];
}

// Bottom-up: p and expansion size of every node of depth `d`; pixels learn when they are born.
__global__ void k_tree_pd(TreeBuild t, int d)
{
  const unsigned base = t.lvl_base[d], cnt = t.lvl_count[d];
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += gridDim.x * blockDim.x) {
    ONode& n = t.nodes[base + i];
    if (n.len == 1) {
      int p = -1;
      if (n.b > n.a) {
        const unsigned long long m = t.omag[n.a];
        p = 63 - __clzll((long long)m);
      }
      n.p = (signed char)p;
      continue;
    }
    if (n.fc == 0) {
      n.p = -1;
      n.d = 0;
      continue;
    }
    ONode& c0 = t.nodes[n.fc];
    ONode& c1 = t.nodes[n.fc + 1];
    const int p = c0.p > c1.p ? c0.p : c1.p;
    unsigned D = 0;
    if (p >= 0) {
      const bool s0 = c0.p == p;
      D = 1 + (s0 ? (c0.len == 1 ? 1u : c0.d) : 0u);
      const bool need1 = s0;
      const bool s1 = !need1 || c1.p == p;
      D += (need1 ? 1u : 0u) + (s1 ? (c1.len == 1 ? 1u : c1.d) : 0u);
      if (c0.len == 1)
        c0.cpl = (signed char)p;
      if (c1.len == 1)
        c1.cpl = (signed char)p;
    }
    n.p = (signed char)p;
    n.d = D;
  }
}

// Collects the pixels the coder creates: key = chunk << 32 | position.
__global__ void k_pix_collect(TreeBuild t, unsigned total_nodes, unsigned long long* keys,
                              unsigned long long* vals, unsigned* npix_total, OutMeta* meta)
{
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < total_nodes;
       i += gridDim.x * blockDim.x) {
    const ONode& n = t.nodes[i];
    if (n.len != 1 || n.cpl < 0)
      continue;
    const unsigned slot = atomicAdd(npix_total, 1u);
    keys[slot] = ((unsigned long long)(unsigned short)n.chunk << 32) | n.start;
    vals[slot] = i;
    atomicAdd(&meta[n.chunk].npix, 1u);
  }
}

__global__ void k_pix_offsets(OutMeta* meta, int nchunks)
{
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    unsigned long long acc = 0;
    for (int c = 0; c < nchunks; c++) {
      meta[c].pix0 = acc;
      acc += (meta[c].npix + 63) & ~63u;  // keep every chunk's sign words separate
    }
  }
}

// Compact LIP / refinement domain in position order (sorted keys).
__global__ void k_pix_fill(TreeBuild t, const OutMeta* meta, const unsigned long long* keys,
                           const unsigned long long* vals, unsigned npix_total,
                           const unsigned char* osign, int8_t* pleaf, int8_t* cmap,
                           unsigned long long* mag, unsigned* signs, const unsigned* chunk_first)
{
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix_total;
       i += gridDim.x * blockDim.x) {
    const unsigned c = unsigned(keys[i] >> 32);
    ONode& n = t.nodes[unsigned(vals[i])];
    const unsigned long long j = meta[c].pix0 + (i - chunk_first[c]);
    pleaf[j] = n.p;
    cmap[j] = n.cpl;
    const bool has = n.b > n.a;
    mag[j] = has ? t.omag[n.a] : 0ull;
    if (has && osign[n.a])
      atomicOr(&signs[j >> 5], 1u << (j & 31));
    n.d = unsigned(j - meta[c].pix0);
  }
}

// first sorted index of every chunk
__global__ void k_pix_first(const unsigned long long* keys, unsigned npix_total, unsigned* chunk_first)
{
  for (unsigned i = blockIdx.x * blockDim.x + threadIdx.x; i < npix_total;
       i += gridDim.x * blockDim.x) {
    const unsigned c = unsigned(keys[i] >> 32);
    if (i == 0 || unsigned(keys[i - 1] >> 32) != c)
      chunk_first[c] = i;
  }
}

// ---------------------------------------------------------------------------------------------
// tree policy
// ---------------------------------------------------------------------------------------------

struct Tree1D {
  static constexpr bool kHasLeaf8 = false;
  static constexpr bool kIsOutlierTree = true;   // profile ranges of this encoder are named enc1d.*
  struct Data {
    const ONode* nodes;
    const OutMeta* meta;
    const unsigned char* osign;
  };
  static __device__ __forceinline__ void pd(const Data& t, const ChunkDev&, unsigned, node_t nd,
                                            int& p, unsigned& d)
  {
    const ONode& n = t.nodes[nd];
    p = n.p;
    d = n.d;
  }
  static __device__ __forceinline__ int children(const Data& t, const ChunkDev&, unsigned c,
                                                 node_t nd, ChildRec* out)
  {
    const ONode& n = t.nodes[nd];
    for (int k = 0; k < 2; k++) {
      const ONode& ch = t.nodes[n.fc + k];
      ChildRec& r = out[k];
      r.id = n.fc + k;
      r.p = ch.p;
      if (ch.len == 1) {
        r.kind = 0;
        r.d = 0;
        r.lis_desc = 0;
        r.sign = ch.b > ch.a ? t.osign[ch.a] : 0u;
      }
      else {
        r.kind = ch.len <= 2 ? 1 : 2;
        r.d = ch.d;
        r.lis_desc = t.meta[c].nlis - 1 - ch.depth;
        r.sign = 0;
      }
    }
    return 2;
  }
  static __device__ __forceinline__ int planes(const Data& t, const ChunkDev&, unsigned c)
  {
    return int(t.nodes[t.meta[c].root].p) + 1;
  }
  static __device__ __forceinline__ int num_roots(const Data&, const ChunkDev&, unsigned) { return 2; }
  static __device__ __forceinline__ void root(const Data& t, const ChunkDev&, unsigned c, int r,
                                              node_t& nd, unsigned& lis_desc, unsigned& order)
  {
    nd = t.nodes[t.meta[c].root].fc + r;
    lis_desc = t.meta[c].nlis - 1 - 1;
    order = unsigned(r);
  }
};

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------

void OutlierCoder::encode(const std::vector<unsigned long long>& total_len, double tol,
                          std::vector<EncResult>& results, cudaStream_t st)
{
  const int nchunks = int(total_len.size());
  results.assign(nchunks, EncResult());
  const size_t total = size_t(ooff[nchunks]);
  if (total == 0)
    return;
  // meta
  std::vector<OutMeta> hm(nchunks);
  unsigned long long cap_nodes = nchunks;
  int max_depth = 1;
  for (int c = 0; c < nchunks; c++) {
    OutMeta& m = hm[c];
    std::memset(&m, 0, sizeof(m));
    m.off0 = ooff[c];
    m.off1 = ooff[c + 1];
    m.total_len = total_len[c];
    const int parts = int(num_of_partitions(total_len[c]));
    m.nlis = unsigned(parts + 1);
    max_depth = std::max(max_depth, parts);
    cap_nodes += 2ull * (parts + 1) * (m.off1 - m.off0) + 4;
  }
  meta_.reserve(sizeof(OutMeta) * nchunks);
  rt::h2d(meta_.p, hm.data(), sizeof(OutMeta) * nchunks, st);
  OutMeta* d_meta = meta_.as<OutMeta>();
  omag_.reserve(total * 8);
  osign_.reserve(total);
  unsigned long long* d_omag = omag_.as<unsigned long long>();
  unsigned char* d_osign = osign_.as<unsigned char>();
  const double* d_err = oerr_.as<double>();
  const unsigned* d_pos = opos_.as<unsigned>();

  LAUNCH(k_out_maxerr, dim3(8, nchunks), dim3(256), 0, st, d_meta, d_err, nchunks);
  LAUNCH(k_out_width, dim3((nchunks + 63) / 64), dim3(64), 0, st, d_meta, nchunks);
  LAUNCH(k_out_quant, dim3(8, nchunks), dim3(256), 0, st, d_meta, d_err, tol, d_omag, d_osign);

  // tree
  nodes_.reserve(sizeof(ONode) * cap_nodes);
  small_.reserve(1024);
  rt::dset(small_.p, 0, 1024, st);
  TreeBuild tb;
  tb.nodes = nodes_.as<ONode>();
  tb.cap = cap_nodes;
  tb.lvl_base = small_.as<unsigned>();
  tb.lvl_count = tb.lvl_base + 64;
  tb.err = tb.lvl_base + 128;
  unsigned* d_npix = tb.lvl_base + 129;
  tb.opos = d_pos;
  tb.omag = d_omag;
  LAUNCH(k_tree_roots, dim3((nchunks + 63) / 64), dim3(64), 0, st, tb, d_meta, nchunks);
  for (int d = 0; d < max_depth; d++) {
    LAUNCH(k_tree_split, dim3(64), dim3(256), 0, st, tb, d);
    LAUNCH(k_tree_next, dim3(1), dim3(32), 0, st, tb, d);
  }
  for (int d = max_depth; d >= 0; d--)
    LAUNCH(k_tree_pd, dim3(64), dim3(256), 0, st, tb, d);

  unsigned lvl[130];
  rt::d2h(lvl, small_.p, sizeof(lvl), st);
  rt::sync(st);
  if (lvl[128])
    throw std::runtime_error("outlier tree overflow");
  const unsigned total_nodes = lvl[max_depth] + lvl[64 + max_depth];

  // pixels in position order
  pkeys_[0].reserve((size_t)total_nodes * 8 + 8); pkeys_[1].reserve((size_t)total_nodes * 8 + 8);
  pvals_[0].reserve((size_t)total_nodes * 8 + 8); pvals_[1].reserve((size_t)total_nodes * 8 + 8);
  unsigned long long* k0 = pkeys_[0].as<unsigned long long>();
  unsigned long long* k1 = pkeys_[1].as<unsigned long long>();
  unsigned long long* v0 = pvals_[0].as<unsigned long long>();
  unsigned long long* v1 = pvals_[1].as<unsigned long long>();
  LAUNCH(k_pix_collect, dim3(64), dim3(256), 0, st, tb, total_nodes, k0, v0, d_npix, d_meta);
  LAUNCH(k_pix_offsets, dim3(1), dim3(32), 0, st, d_meta, nchunks);
  unsigned npix_total = 0;
  rt::d2h(&npix_total, d_npix, 4, st);
  rt::d2h(hm.data(), d_meta, sizeof(OutMeta) * nchunks, st);
  rt::sync(st);
  for (int c = 0; c < nchunks; c++)
    if (hm[c].off1 > hm[c].off0 && hm[c].width == 0)
      throw std::runtime_error("FE_INVALID in outlier coder");
  const size_t sort_bytes = sort_tmp_bytes(npix_total);
  sort_tmp_.reserve(sort_bytes);
  sort_pairs_u64(k0, k1, v0, v1, npix_total, 48, sort_tmp_.p, sort_bytes, st);
  size_t pix_cap = 64;
  for (int c = 0; c < nchunks; c++)
    pix_cap += (hm[c].npix + 63) & ~63u;
  ppleaf_.reserve(pix_cap); pcmap_.reserve(pix_cap);
  pmag_.reserve(pix_cap * 8);
  psigns_.reserve(pix_cap / 8 + 64);
  rt::dset(psigns_.p, 0, pix_cap / 8 + 64, st);
  rt::dset(pcmap_.p, 0xFF, pix_cap, st);
  rt::dset(ppleaf_.p, 0xFF, pix_cap, st);
  first_.reserve(4 * (nchunks + 1));
  rt::dset(first_.p, 0, 4 * (nchunks + 1), st);
  unsigned* d_first = first_.as<unsigned>();
  int8_t* d_pleaf = ppleaf_.as<int8_t>();
  int8_t* d_cmap = pcmap_.as<int8_t>();
  unsigned long long* d_pmag = pmag_.as<unsigned long long>();
  unsigned* d_psigns = psigns_.as<unsigned>();
  if (npix_total) {
    LAUNCH(k_pix_first, dim3(64), dim3(256), 0, st, k1, npix_total, d_first);
    LAUNCH(k_pix_fill, dim3(64), dim3(256), 0, st, tb, d_meta, k1, v1, npix_total, d_osign, d_pleaf,
           d_cmap, d_pmag, d_psigns, d_first);
  }

  // chunk array of the coder: the compact pixel arrays are its LIP / refinement domain
  std::vector<ChunkDev> oc(nchunks);
  size_t max_n = 0;
  for (int c = 0; c < nchunks; c++) {
    ChunkDev& d = oc[c];
    std::memset(&d, 0, sizeof(d));
    d.n = hm[c].npix;
    d.nx = hm[c].npix; d.ny = d.nz = 1;
    d.is_const = (hm[c].off1 == hm[c].off0) ? 1 : 0;
    d.wide = 1;
    d.budget = ~0ull;
    d.pleaf = d_pleaf + hm[c].pix0;
    d.cmap = d_cmap + hm[c].pix0;
    d.mag = d_pmag + hm[c].pix0;
    d.signs = d_psigns + hm[c].pix0 / 32;
    max_n = std::max<size_t>(max_n, d.n);
  }
  ochunks_.reserve(sizeof(ChunkDev) * nchunks);
  rt::h2d(ochunks_.p, oc.data(), sizeof(ChunkDev) * nchunks, st);
  rt::sync(st);
  Tree1D::Data tree{tb.nodes, d_meta, d_osign};
  std::vector<unsigned long long> nodes_of(nchunks);
  for (int c = 0; c < nchunks; c++)
    nodes_of[c] = 2ull * (hm[c].nlis) * (hm[c].off1 - hm[c].off0) + 8;
  auto bound = [&](int c, const ChunkDev& hc) {
    return (nodes_of[c] + hc.n) * (unsigned long long)(hc.planes + 1) + 64;
  };
  run_encoder<Tree1D>(work_, ochunks_.as<ChunkDev>(), nchunks, max_n, tree, total_nodes, max_depth + 1,
                      bound, results, st);
}

}  // namespace sperr_b200
