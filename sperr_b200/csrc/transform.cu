// Conditioner, CDF 9/7 lifting transform and mid-tread quantiser kernels.
//
// Arithmetic follows the reference built with -ffp-contract=off ("STRICT" flavour): every fp64
// multiply / add is a separately rounded __dmul_rn / __dadd_rn (and the file is compiled with
// -fmad=false), in the reference's operation order:
//   Conditioner::condition / m_calc_mean     /root/reference/src/Conditioner.cpp:10-64,119-135
//   QccWAVCDF97AnalysisSymmetric / Synthesis /root/reference/src/CDF97.cpp:598-666
//   m_dwt3d_one_level / m_idwt3d_one_level   /root/reference/src/CDF97.cpp:387-474
//   m_midtread_quantize / inv_quantize       /root/reference/src/SPECK_FLT.cpp:311-399
//   m_estimate_mse_midtread                  /root/reference/src/SPECK_FLT.cpp:237-266
#include "kernels.h"

namespace sperr_b200 {

// ---------------------------------------------------------------------------------------------
// source access
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ double load_src(const SrcVol& s, unsigned long long idx)
{
  return s.is_float ? double(reinterpret_cast<const float*>(s.ptr)[idx])
                    : reinterpret_cast<const double*>(s.ptr)[idx];
}

__device__ __forceinline__ unsigned long long src_index(const SrcVol& s, const ChunkDev& c,
                                                        unsigned x, unsigned y, unsigned z)
{
  return (unsigned long long)(c.z0 + z) * s.vx * s.vy + (unsigned long long)(c.y0 + y) * s.vx +
         (c.x0 + x);
}

// Order-preserving map double -> uint64 (for atomic min / max).
__device__ __forceinline__ unsigned long long dkey(double v)
{
  unsigned long long b = (unsigned long long)__double_as_longlong(v);
  return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}

// ---------------------------------------------------------------------------------------------
// conditioner: strided mean in the reference's association order, constant test, min / max
// ---------------------------------------------------------------------------------------------

// Strides made of whole rows of a float volume can be staged through shared memory with coalesced
// 128-bit loads (k_stride_stats_rows); everything else is walked by k_stride_stats.
constexpr int kStatStrides = 128;   // strides (= threads) of a block
constexpr int kStatSeg = 64;        // floats of a row staged per step and stride
__device__ __forceinline__ bool stats_by_rows(const SrcVol& src, const ChunkDev& ch, unsigned ns)
{
  return src.is_float && ns != 0 && ch.nx % kStatSeg == 0 && (ch.n / ns) % ch.nx == 0 && ch.n % ns == 0 &&
         ch.x0 % 4 == 0 && src.vx % 4 == 0 && (reinterpret_cast<unsigned long long>(src.ptr) & 15ull) == 0;
}

// One thread per stride, as below (the additions of a stride happen left to right in one thread,
// src/Conditioner.cpp:119-135), but the values arrive through a [stride][segment] tile that the
// block fills with coalesced loads: 128 strides x 64 floats per step.
__global__ void __launch_bounds__(kStatStrides) k_stride_stats_rows(SrcVol src, ChunkDev* chunks, double* stride_mean,
                                                                 int max_strides, const unsigned* nstrides,
                                                                 unsigned* not_const, int want_minmax)
{
  __shared__ float tile[kStatStrides][kStatSeg + 1];
  const unsigned c = blockIdx.y;
  ChunkDev& ch = chunks[c];
  const unsigned ns = nstrides[c];
  const unsigned s0 = blockIdx.x * kStatStrides;
  if (s0 >= ns || !stats_by_rows(src, ch, ns))
    return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long len = ch.n / ns;
  const unsigned rows = unsigned(len / ch.nx), segs = ch.nx / kStatSeg, ny = ch.ny;
  const unsigned s = s0 + tid;
  const bool active = s < ns;
  const float* const vf = reinterpret_cast<const float*>(src.ptr);
  const double v0 = double(vf[src_index(src, ch, 0, 0, 0)]);
  double acc = 0.0, mn = v0, mx = v0;
  bool diff = false;
  for (unsigned r = 0; r < rows; r++) {
    for (unsigned g = 0; g < segs; g++) {
      // a warp fetches the segments of two strides at a time (16 lanes x 16 bytes each); the loads
      // of eight such pieces are in flight before the first one is stored
      constexpr int kPieces = kStatStrides / 8;   // per lane and step
#pragma unroll
      for (int h = 0; h < kPieces; h += 8) {
        float4 v[8];
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned k = 2 * warp + (lane >> 4) + 8 * (h + j);
          const unsigned ss = s0 + k;
          v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ss < ns) {
            const unsigned rho = ss * rows + r;   // row of the chunk (< ny * nz: 32-bit arithmetic)
            const unsigned z = rho / ny, y = rho - z * ny;
            v[j] = __ldg(reinterpret_cast<const float4*>(vf + src_index(src, ch, g * kStatSeg, y, z)) + (lane & 15));
          }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const unsigned k = 2 * warp + (lane >> 4) + 8 * (h + j);
          float* const t = &tile[k][4 * (lane & 15)];
          t[0] = v[j].x; t[1] = v[j].y; t[2] = v[j].z; t[3] = v[j].w;
        }
      }
      __syncthreads();
      if (active) {
#pragma unroll 8
        for (int i = 0; i < kStatSeg; i++) {
          const double v = double(tile[tid][i]);
          acc = __dadd_rn(acc, v);
          diff |= !(v == v0);
          if (want_minmax) {
            mn = v < mn ? v : mn;
            mx = v > mx ? v : mx;
          }
        }
      }
      __syncthreads();
    }
  }
  if (!active)
    return;
  stride_mean[(size_t)c * max_strides + s] = __ddiv_rn(acc, double(len));
  if (diff)
    atomicOr(&not_const[c], 1u);
  if (want_minmax) {
    atomicMin(&ch.min_key, dkey(mn));
    atomicMax(&ch.max_key, dkey(mx));
  }
  if (s == 0)
    ch.first_val = v0;
}

// One thread per stride; a stride is n / nstrides consecutive values of the chunk (x fastest).
__global__ void k_stride_stats(SrcVol src, ChunkDev* chunks, double* stride_mean, int max_strides,
                               const unsigned* nstrides, unsigned* not_const, int want_minmax)
{
  const unsigned c = blockIdx.y;
  ChunkDev& ch = chunks[c];
  const unsigned ns = nstrides[c];
  const unsigned s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= ns || stats_by_rows(src, ch, ns))
    return;
  const unsigned long long len = ch.n / ns;
  unsigned long long e = (unsigned long long)s * len;
  unsigned x = unsigned(e % ch.nx);
  unsigned y = unsigned((e / ch.nx) % ch.ny);
  unsigned z = unsigned(e / ((unsigned long long)ch.nx * ch.ny));
  const double v0 = load_src(src, src_index(src, ch, 0, 0, 0));
  double acc = 0.0;
  bool diff = false;
  double mn = v0, mx = v0;
  unsigned long long row = src_index(src, ch, 0, y, z);
  for (unsigned long long i = 0; i < len; i++) {
    const double v = load_src(src, row + x);
    acc = __dadd_rn(acc, v);
    diff |= !(v == v0);
    if (want_minmax) {
      mn = v < mn ? v : mn;
      mx = v > mx ? v : mx;
    }
    if (++x == ch.nx) {
      x = 0;
      if (++y == ch.ny) {
        y = 0;
        ++z;
      }
      row = src_index(src, ch, 0, y, z);
    }
  }
  stride_mean[(size_t)c * max_strides + s] = __ddiv_rn(acc, double(len));
  if (diff)
    atomicOr(&not_const[c], 1u);
  if (want_minmax) {
    atomicMin(&ch.min_key, dkey(mn));
    atomicMax(&ch.max_key, dkey(mx));
  }
  if (s == 0)
    ch.first_val = v0;
}

__global__ void k_mean_finish(ChunkDev* chunks, const double* stride_mean, int max_strides,
                              const unsigned* nstrides, const unsigned* not_const, int nchunks)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks)
    return;
  ChunkDev& ch = chunks[c];
  const unsigned ns = nstrides[c];
  double sum = 0.0;
  for (unsigned s = 0; s < ns; s++)
    sum = __dadd_rn(sum, stride_mean[(size_t)c * max_strides + s]);
  ch.mean = __ddiv_rn(sum, double(ns));
  ch.is_const = not_const[c] ? 0 : 1;
}

// coef[i] = double(src) - mean, chunk-linear order.
__global__ void k_gather_sub(SrcVol src, const ChunkDev* chunks)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  if (ch.is_const || ch.fused)   // fused chunks are read straight from the volume by k_fwd3d
    return;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const double mean = ch.mean;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ch.n;
       e += stride) {
    const unsigned x = unsigned(e % ch.nx);
    const unsigned y = unsigned((e / ch.nx) % ch.ny);
    const unsigned z = unsigned(e / ((unsigned long long)ch.nx * ch.ny));
    ch.coef[e] = __dsub_rn(load_src(src, src_index(src, ch, x, y, z)), mean);
  }
}

// ---------------------------------------------------------------------------------------------
// lifting steps on a line held in shared memory as evens (el) followed by odds (ol).
// `get/put` index the line; `tid/nthr` enumerate the cooperating threads; `sync` separates steps.
// ---------------------------------------------------------------------------------------------

template <bool FMA, typename Line, typename Sync>
__device__ __forceinline__ void lift_forward(const CdfC& k, Line L, int len, int tid, int nthr,
                                             Sync sync)
{
  const int el = len - len / 2, ol = len / 2;
  for (int i = tid; i < ol; i += nthr) {
    const int i1 = i + 1 < el ? i + 1 : el - 1;
    L.o(i) = lift_add<FMA>(L.o(i), k.ALPHA, __dadd_rn(L.e(i), L.e(i1)));
  }
  sync();
  for (int i = tid; i < el; i += nthr) {
    const int a = i > 0 ? i - 1 : 0, b = i < ol ? i : ol - 1;
    L.e(i) = lift_add<FMA>(L.e(i), k.BETA, __dadd_rn(L.o(a), L.o(b)));
  }
  sync();
  for (int i = tid; i < ol; i += nthr) {
    const int i1 = i + 1 < el ? i + 1 : el - 1;
    L.o(i) = lift_add<FMA>(L.o(i), k.GAMMA, __dadd_rn(L.e(i), L.e(i1)));
  }
  sync();
  for (int i = tid; i < el; i += nthr) {
    const int a = i > 0 ? i - 1 : 0, b = i < ol ? i : ol - 1;
    L.e(i) = lift_scale_fwd<FMA>(k, L.e(i), __dadd_rn(L.o(a), L.o(b)));
  }
  sync();
  for (int i = tid; i < ol; i += nthr)
    L.o(i) = __dmul_rn(L.o(i), -k.INV_EPSILON);
  sync();
}

template <bool FMA, typename Line, typename Sync>
__device__ __forceinline__ void lift_inverse(const CdfC& k, Line L, int len, int tid, int nthr,
                                             Sync sync)
{
  const int el = len - len / 2, ol = len / 2;
  for (int i = tid; i < ol; i += nthr)
    L.o(i) = __dmul_rn(L.o(i), -k.EPSILON);
  sync();
  for (int i = tid; i < el; i += nthr) {
    const int a = i > 0 ? i - 1 : 0, b = i < ol ? i : ol - 1;
    L.e(i) = lift_scale_inv<FMA>(k, L.e(i), __dadd_rn(L.o(a), L.o(b)));
  }
  sync();
  for (int i = tid; i < ol; i += nthr) {
    const int i1 = i + 1 < el ? i + 1 : el - 1;
    L.o(i) = lift_sub<FMA>(L.o(i), k.GAMMA, __dadd_rn(L.e(i), L.e(i1)));
  }
  sync();
  for (int i = tid; i < el; i += nthr) {
    const int a = i > 0 ? i - 1 : 0, b = i < ol ? i : ol - 1;
    L.e(i) = lift_sub<FMA>(L.e(i), k.BETA, __dadd_rn(L.o(a), L.o(b)));
  }
  sync();
  for (int i = tid; i < ol; i += nthr) {
    const int i1 = i + 1 < el ? i + 1 : el - 1;
    L.o(i) = lift_sub<FMA>(L.o(i), k.ALPHA, __dadd_rn(L.e(i), L.e(i1)));
  }
  sync();
}

struct SmemLine {  // contiguous line: evens then odds
  double* s;
  int el;
  __device__ __forceinline__ double& e(int i) const { return s[i]; }
  __device__ __forceinline__ double& o(int i) const { return s[el + i]; }
};

struct SmemCol {  // column `cx` of a [len][tw] tile
  double* s;
  int el, tw, cx;
  __device__ __forceinline__ double& e(int i) const { return s[i * tw + cx]; }
  __device__ __forceinline__ double& o(int i) const { return s[(el + i) * tw + cx]; }
};

// X pass: one warp per row of the level box (lx, ly, lz); rows are contiguous in memory.
template <bool INVERSE>
__global__ void k_dwt_x(const ChunkDev* chunks, const int* ids, CdfC k, int lx, int ly, int lz)
{
  DYN_SMEM(double, smem);
  const ChunkDev& ch = chunks[ids[blockIdx.y]];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  const long long row = (long long)blockIdx.x * wpb + warp;
  const bool live = row < (long long)ly * lz;   // whole warps are live or not
  double* s = smem + (size_t)warp * lx;
  const int el = lx - lx / 2;
  double* g = nullptr;
  if (live) {
    const int y = int(row % ly), z = int(row / ly);
    g = ch.coef + ((size_t)z * ch.ny + y) * ch.nx;
    if (!INVERSE) {
      for (int i = lane; i < lx; i += 32)
        s[(i & 1) ? el + (i >> 1) : (i >> 1)] = g[i];
    }
    else {
      for (int i = lane; i < lx; i += 32)
        s[i] = g[i];
    }
  }
  __syncwarp();
  if (live) {
    SmemLine L{s, el};
    auto sync = [] { __syncwarp(); };
    if (!INVERSE) {
      if (k.fma)
        lift_forward<true>(k, L, lx, lane, 32, sync);
      else
        lift_forward<false>(k, L, lx, lane, 32, sync);
      for (int i = lane; i < lx; i += 32)
        g[i] = s[i];
    }
    else {
      if (k.fma)
        lift_inverse<true>(k, L, lx, lane, 32, sync);
      else
        lift_inverse<false>(k, L, lx, lane, 32, sync);
      for (int i = lane; i < lx; i += 32)
        g[i] = s[(i & 1) ? el + (i >> 1) : (i >> 1)];
    }
  }
}

// Strided pass (Y or Z): a block owns `tw` neighbouring x positions and the full line.
//   axis == 1: lines run along y at fixed z  (stride nx),   outer index = z in [0, lo)
//   axis == 2: lines run along z at fixed y  (stride nx*ny), outer index = y in [0, lo)
template <bool INVERSE>
__global__ void k_dwt_col(const ChunkDev* chunks, const int* ids, CdfC k, int axis, int lx, int len,
                          int lo, int tw)
{
  DYN_SMEM(double, smem);
  const ChunkDev& ch = chunks[ids[blockIdx.y]];
  const int tiles_x = (lx + tw - 1) / tw;
  const int tile = blockIdx.x % tiles_x, outer = blockIdx.x / tiles_x;
  const int cx = threadIdx.x % tw, r0 = threadIdx.x / tw, th = blockDim.x / tw;
  const int x = tile * tw + cx;
  const bool in = x < lx;
  const size_t plane = (size_t)ch.nx * ch.ny;
  const size_t stride = axis == 1 ? ch.nx : plane;
  double* g = ch.coef + (axis == 1 ? (size_t)outer * plane : (size_t)outer * ch.nx) + x;
  const int el = len - len / 2;
  (void)lo;
  if (in) {
    if (!INVERSE)
      for (int i = r0; i < len; i += th)
        smem[((i & 1) ? el + (i >> 1) : (i >> 1)) * tw + cx] = g[(size_t)i * stride];
    else
      for (int i = r0; i < len; i += th)
        smem[i * tw + cx] = g[(size_t)i * stride];
  }
  __syncthreads();
  SmemCol L{smem, el, tw, cx};
  auto sync = [] { __syncthreads(); };
  // out-of-range columns still walk the barriers; they touch only their own (unused) column
  if (!INVERSE) {
    if (k.fma)
      lift_forward<true>(k, L, in ? len : 0, r0, th, sync);
    else
      lift_forward<false>(k, L, in ? len : 0, r0, th, sync);
  }
  else {
    if (k.fma)
      lift_inverse<true>(k, L, in ? len : 0, r0, th, sync);
    else
      lift_inverse<false>(k, L, in ? len : 0, r0, th, sync);
  }
  if (in) {
    if (!INVERSE)
      for (int i = r0; i < len; i += th)
        g[(size_t)i * stride] = smem[i * tw + cx];
    else
      for (int i = r0; i < len; i += th)
        g[(size_t)i * stride] = smem[((i & 1) ? el + (i >> 1) : (i >> 1)) * tw + cx];
  }
}

// ---------------------------------------------------------------------------------------------
// quantiser
// ---------------------------------------------------------------------------------------------

__global__ void k_absmax(ChunkDev* chunks)
{
  ChunkDev& ch = chunks[blockIdx.y];
  if (ch.is_const || ch.fused)   // k_fwd3d tracks the maximum itself
    return;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long m = 0;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ch.n;
       e += stride) {
    const unsigned long long b =
        (unsigned long long)__double_as_longlong(ch.coef[e]) & 0x7fffffffffffffffull;
    m = b > m ? b : m;
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, m, o);
    m = t > m ? t : m;
  }
  if ((threadIdx.x & 31) == 0 && m)
    atomicMax(&ch.max_bits, m);
}

// Integer width selection (src/SPECK_FLT.cpp:318-337): llrint(max|v| / q) with FE_INVALID check.
__global__ void k_qdecide(ChunkDev* chunks, int nchunks)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= nchunks)
    return;
  ChunkDev& ch = chunks[c];
  if (ch.is_const)
    return;
  const double mx = __longlong_as_double((long long)ch.max_bits);
  const double r = __ddiv_rn(mx, ch.q);
  // NaN, Inf or beyond the long long range raise FE_INVALID in llrint
  if (!(r < 9223372036854775807.0)) {
    ch.fe_invalid = 1;
    return;
  }
  const long long maxll = __double2ll_rn(r);
  ch.wide = maxll > 0xFFFFFFFFll ? 1 : 0;
}

__global__ void k_quantize(const ChunkDev* chunks)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  if (ch.is_const || ch.fe_invalid)
    return;
  const double inv = __ddiv_rn(1.0, ch.q);
  const unsigned long long n32 = (ch.n + 31) & ~31ull;  // whole warps
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n32;
       e += stride) {
    long long ll = 0;
    if (e < ch.n)
      ll = __double2ll_rn(__dmul_rn(ch.coef[e], inv));
    const unsigned long long m = (unsigned long long)(ll < 0 ? -ll : ll);
    const unsigned sb = __ballot_sync(0xffffffffu, ll >= 0);
    if (e < ch.n) {
      if (ch.wide)
        reinterpret_cast<unsigned long long*>(ch.mag)[e] = m;
      else
        reinterpret_cast<unsigned*>(ch.mag)[e] = unsigned(m);
      ch.pleaf[e] = int8_t(63 - __clzll((long long)m));
    }
    if ((threadIdx.x & 31) == 0)
      ch.signs[e >> 5] = sb;
  }
}

__global__ void k_inv_quantize(const ChunkDev* chunks)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  if (ch.is_const)
    return;
  const double q = ch.q;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ch.n;
       e += stride) {
    const double m = ch.wide ? __ull2double_rn(reinterpret_cast<const unsigned long long*>(ch.mag)[e])
                             : __uint2double_rn(reinterpret_cast<const unsigned*>(ch.mag)[e]);
    const bool pos = (ch.signs[e >> 5] >> (e & 31)) & 1u;
    ch.coef[e] = __dmul_rn(__dmul_rn(q, m), pos ? 1.0 : -1.0);
  }
}

// PSNR mode: strided sums of fma(-q, rint(v / q), v)^2, 4096 values per stride plus a tail stride.
__global__ void k_mse_strides(const ChunkDev* chunks, const int* ids, const double* qs,
                              double* partial, int max_strides)
{
  const int slot = blockIdx.y;
  const ChunkDev& ch = chunks[ids[slot]];
  const unsigned long long ns = ch.n / 4096;  // full strides; stride `ns` is the tail
  const unsigned long long s = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s > ns)
    return;
  const double q = qs[slot];
  const double rcp = __ddiv_rn(1.0, q);
  const unsigned long long beg = s * 4096, end = s == ns ? ch.n : beg + 4096;
  double acc = 0.0;
  for (unsigned long long i = beg; i < end; i++) {
    const double v = ch.coef[i];
    const double d = __fma_rn(-q, rint(__dmul_rn(v, rcp)), v);
    acc = __dadd_rn(acc, __dmul_rn(d, d));
  }
  partial[(size_t)slot * max_strides + s] = acc;
}

__global__ void k_mse_finish(const ChunkDev* chunks, const int* ids, const double* partial,
                             int max_strides, double* mse, int nslots)
{
  const int slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= nslots)
    return;
  const ChunkDev& ch = chunks[ids[slot]];
  const unsigned long long ns = ch.n / 4096;
  double total = 0.0;
  for (unsigned long long s = 0; s <= ns; s++)
    total = __dadd_rn(total, partial[(size_t)slot * max_strides + s]);
  mse[slot] = __ddiv_rn(total, double(ch.n));
}

// ---------------------------------------------------------------------------------------------
// decode side: add mean / constant fill and scatter into the output volume
// ---------------------------------------------------------------------------------------------

__global__ void k_scatter_out(SrcVol dst, const ChunkDev* chunks)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  if (ch.fused && !ch.is_const)   // written by the fused inverse transform
    return;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < ch.n;
       e += stride) {
    const unsigned x = unsigned(e % ch.nx);
    const unsigned y = unsigned((e / ch.nx) % ch.ny);
    const unsigned z = unsigned(e / ((unsigned long long)ch.nx * ch.ny));
    const double v = ch.is_const ? ch.first_val : __dadd_rn(ch.coef[e], ch.mean);
    const unsigned long long o = src_index(dst, ch, x, y, z);
    if (dst.is_float)
      reinterpret_cast<float*>(const_cast<void*>(dst.ptr))[o] = __double2float_rn(v);
    else
      reinterpret_cast<double*>(const_cast<void*>(dst.ptr))[o] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------

int& fma_flavour()
{
  static int f = std::getenv("SPERR_B200_FMA") ? std::atoi(std::getenv("SPERR_B200_FMA")) : 0;
  return f;
}

CdfC cdf_constants()
{
  // Same expressions as include/CDF97.h:136-147, evaluated in plain IEEE double arithmetic
  // (volatile blocks compile-time contraction or reassociation).
  volatile double h0 = 0.602949018236, h1 = 0.266864118443, h2 = -0.078223266529,
                  h3 = -0.016864118443, h4 = 0.026748757411;
  volatile double t1 = 2.0 * h4;
  volatile double t2 = t1 * h1;
  volatile double t3 = t2 / h3;
  volatile double r0 = h0 - t3;
  volatile double u1 = h2 - h4;
  volatile double u2 = h4 * h1;
  volatile double u3 = u2 / h3;
  volatile double r1 = u1 - u3;
  volatile double v1 = h1 - h3;
  volatile double v2 = h3 * r0;
  volatile double v3 = v2 / r1;
  volatile double s0 = v1 - v3;
  volatile double w1 = h2 - h4;
  volatile double w2 = 2.0 * w1;
  volatile double t0 = h0 - w2;
  CdfC c;
  c.ALPHA = h4 / h3;
  c.BETA = h3 / r1;
  c.GAMMA = r1 / s0;
  c.DELTA = s0 / t0;
  volatile double sq = sqrt(2.0);
  c.EPSILON = sq * t0;
  volatile double eps = c.EPSILON;
  c.INV_EPSILON = 1.0 / eps;
  c.fma = fma_flavour() ? 1 : 0;
  return c;
}

void launch_stats(const SrcVol& src, ChunkDev* d_chunks, int nchunks, double* d_stride_mean,
                  int max_strides, const unsigned* d_nstrides, unsigned* d_not_const,
                  bool want_minmax, cudaStream_t st)
{
  dim3 grid((max_strides + 127) / 128, nchunks);
  if (src.is_float)   // chunks whose strides are whole rows (every kernel checks per chunk)
    LAUNCH(k_stride_stats_rows, grid, dim3(kStatStrides), 0, st, src, d_chunks, d_stride_mean, max_strides,
           d_nstrides, d_not_const, want_minmax ? 1 : 0);
  LAUNCH(k_stride_stats, grid, dim3(128), 0, st, src, d_chunks, d_stride_mean, max_strides,
         d_nstrides, d_not_const, want_minmax ? 1 : 0);
  LAUNCH(k_mean_finish, dim3((nchunks + 63) / 64), dim3(64), 0, st, d_chunks, d_stride_mean,
         max_strides, d_nstrides, d_not_const, nchunks);
}

void launch_gather(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                   cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 2048));
  LAUNCH(k_gather_sub, dim3(gx, nchunks), dim3(256), 0, st, src, d_chunks);
}

static int pick_tw(int len)
{
  // [len][tw] fp64 tile must fit in shared memory (we allow up to ~200 KB)
  int tw = 32;
  while (tw > 1 && (size_t)len * tw * 8 > 200 * 1024)
    tw >>= 1;
  return tw;
}

static void ensure_smem_attr()
{
#ifndef SPERR_EMUL
  static rt::OncePerDevice once;
  if (once.first()) {
    const int big = 220 * 1024;
    RT_CHECK(cudaFuncSetAttribute(k_dwt_col<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    RT_CHECK(cudaFuncSetAttribute(k_dwt_col<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    RT_CHECK(cudaFuncSetAttribute(k_dwt_x<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
    RT_CHECK(cudaFuncSetAttribute(k_dwt_x<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big));
  }
#endif
}

static void pass_x(bool inverse, const ChunkDev* d_chunks, const int* d_ids, int nids, int lx,
                   int ly, int lz, const CdfC& k, cudaStream_t st)
{
  if ((size_t)lx * 8 > 200 * 1024)
    throw std::runtime_error("chunk x extent too large for the line kernel");
  int wpb = 8;
  while (wpb > 1 && (size_t)wpb * lx * 8 > 200 * 1024)
    wpb >>= 1;
  const long long rows = (long long)ly * lz;
  dim3 grid(unsigned((rows + wpb - 1) / wpb), nids);
  const size_t smem = (size_t)wpb * lx * 8;
  if (inverse)
    LAUNCH(k_dwt_x<true>, grid, dim3(wpb * 32), smem, st, d_chunks, d_ids, k, lx, ly, lz);
  else
    LAUNCH(k_dwt_x<false>, grid, dim3(wpb * 32), smem, st, d_chunks, d_ids, k, lx, ly, lz);
}

static void pass_col(bool inverse, int axis, const ChunkDev* d_chunks, const int* d_ids, int nids,
                     int lx, int len, int lo, const CdfC& k, cudaStream_t st)
{
  const int tw = pick_tw(len);
  if ((size_t)len * tw * 8 > 200 * 1024)
    throw std::runtime_error("chunk extent too large for the column kernel");
  const int tiles_x = (lx + tw - 1) / tw;
  dim3 grid(unsigned(tiles_x) * unsigned(lo), nids);
  const size_t smem = (size_t)len * tw * 8;
  if (inverse)
    LAUNCH(k_dwt_col<true>, grid, dim3(256), smem, st, d_chunks, d_ids, k, axis, lx, len, lo, tw);
  else
    LAUNCH(k_dwt_col<false>, grid, dim3(256), smem, st, d_chunks, d_ids, k, axis, lx, len, lo, tw);
}

// Full multi-level transform of every chunk in `d_ids` (all of shape nx, ny, nz).
// is_2d: single-plane chunks transformed with dwt2d (CDF97::dwt2d, src/CDF97.cpp:102-106).
void launch_dwt(bool inverse, const ChunkDev* d_chunks, const int* d_ids, int nids, uint32_t nx,
                uint32_t ny, uint32_t nz, bool is_2d, cudaStream_t st)
{
  ensure_smem_attr();
  const CdfC k = cdf_constants();
  const int dy = is_2d ? -1 : can_use_dyadic(nx, ny, nz);
  if (dy >= 0) {
    for (int s = 0; s < dy; s++) {
      const int lev = inverse ? dy - 1 - s : s;
      const int lx = int(calc_approx_detail_len(nx, lev)[0]);
      const int ly = int(calc_approx_detail_len(ny, lev)[0]);
      const int lz = int(calc_approx_detail_len(nz, lev)[0]);
      if (!inverse) {
        pass_x(false, d_chunks, d_ids, nids, lx, ly, lz, k, st);
        pass_col(false, 1, d_chunks, d_ids, nids, lx, ly, lz, k, st);
        pass_col(false, 2, d_chunks, d_ids, nids, lx, lz, ly, k, st);
      }
      else {
        pass_col(true, 2, d_chunks, d_ids, nids, lx, lz, ly, k, st);
        pass_col(true, 1, d_chunks, d_ids, nids, lx, ly, lz, k, st);
        pass_x(true, d_chunks, d_ids, nids, lx, ly, lz, k, st);
      }
    }
    return;
  }
  // wavelet-packet: all z levels on full (x, y) extents, then per-plane 2D levels
  const int nzx = is_2d ? 0 : int(num_of_xforms(nz));
  const int nxy = int(num_of_xforms(std::min(nx, ny)));
  if (!inverse) {
    for (int lev = 0; lev < nzx; lev++)
      pass_col(false, 2, d_chunks, d_ids, nids, nx, int(calc_approx_detail_len(nz, lev)[0]), ny, k, st);
    for (int lev = 0; lev < nxy; lev++) {
      const int lx = int(calc_approx_detail_len(nx, lev)[0]);
      const int ly = int(calc_approx_detail_len(ny, lev)[0]);
      pass_x(false, d_chunks, d_ids, nids, lx, ly, nz, k, st);
      pass_col(false, 1, d_chunks, d_ids, nids, lx, ly, nz, k, st);
    }
  }
  else {
    for (int lev = nxy - 1; lev >= 0; lev--) {
      const int lx = int(calc_approx_detail_len(nx, lev)[0]);
      const int ly = int(calc_approx_detail_len(ny, lev)[0]);
      pass_col(true, 1, d_chunks, d_ids, nids, lx, ly, nz, k, st);
      pass_x(true, d_chunks, d_ids, nids, lx, ly, nz, k, st);
    }
    for (int lev = nzx - 1; lev >= 0; lev--)
      pass_col(true, 2, d_chunks, d_ids, nids, nx, int(calc_approx_detail_len(nz, lev)[0]), ny, k, st);
  }
}

void launch_absmax(ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 1024));
  LAUNCH(k_absmax, dim3(gx, nchunks), dim3(256), 0, st, d_chunks);
}

void launch_qdecide(ChunkDev* d_chunks, int nchunks, cudaStream_t st)
{
  LAUNCH(k_qdecide, dim3((nchunks + 63) / 64), dim3(64), 0, st, d_chunks, nchunks);
}

void launch_quantize(const ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 2048));
  LAUNCH(k_quantize, dim3(gx, nchunks), dim3(256), 0, st, d_chunks);
}

void launch_inv_quantize(const ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 2048));
  LAUNCH(k_inv_quantize, dim3(gx, nchunks), dim3(256), 0, st, d_chunks);
}

void launch_mse(const ChunkDev* d_chunks, const int* d_ids, const double* d_qs, double* d_partial,
                int max_strides, double* d_mse, int nslots, cudaStream_t st)
{
  LAUNCH(k_mse_strides, dim3((max_strides + 127) / 128, nslots), dim3(128), 0, st, d_chunks, d_ids,
         d_qs, d_partial, max_strides);
  LAUNCH(k_mse_finish, dim3((nslots + 63) / 64), dim3(64), 0, st, d_chunks, d_ids, d_partial,
         max_strides, d_mse, nslots);
}

void launch_scatter_out(const SrcVol& dst, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                        cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n + 255) / 256, 2048));
  LAUNCH(k_scatter_out, dim3(gx, nchunks), dim3(256), 0, st, dst, d_chunks);
}

}  // namespace sperr_b200
