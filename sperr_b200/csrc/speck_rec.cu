// Magnitude reconstruction after the SPECK sorting passes have been decoded (speck_dec*.cuh).
//
// Reference behaviour (bit-exact): SPECK_INT<T>::m_refinement_pass_decode and the initial value of a
// newly significant coefficient (/root/reference/src/SPECK_INT.cpp:359-469), followed by
//   mode 0: SPECK_FLT::m_midtread_inv_quantize (/root/reference/src/SPECK_FLT.cpp:373-399)
//   mode 1: Outlier_Coder::m_inverse_quantize (src/Outlier_Coder.cpp:206-234) applied as in
//           src/SPECK_FLT.cpp:576-585.
// How: the reference refines plane by plane, walking the LSP mask each time. Here every coefficient
// gathers its own bits: in plane n it owns bit number rank_n(i) of that plane's refinement section,
// where rank_n(i) counts the coefficients before i (raster order) that were significant before
// plane n. Ranks are per-block counts (k_rec_count), scanned over blocks (k_rec_scan), plus a
// ballot prefix inside the block (k_rec_apply) -- one grid-wide pass over all chunks of the batch
// instead of one pass per plane inside the per-chunk decoder.
#include "speck_dec.h"

namespace sperr_b200 {

constexpr int kRecBlock = 1024;

__device__ __forceinline__ int rec_plane(unsigned v) { return v == 0xFFu ? -1 : int(v & 63u); }

// A CUDA block takes kRecUnits consecutive 1024-coefficient units: their loads are in flight
// together and units without any significant coefficient (most of a sparse outlier array, the fine
// sub-bands of a smooth field) cost no barrier at all.
constexpr int kRecUnits = 8;

// k_rec_apply is bound by its instruction count (~500 per thread and unit on the full path). With
// SPERR_REC_PREFIX the per-plane ranks of every warp's first coefficient (unit base + counts of the
// warps before) are computed once per unit, one warp per plane, instead of by every warp in every
// plane iteration: one barrier more per unit, a REDUX, a select and a global load less per plane
// iteration. Measured on B200 (1024^3, two runs): d.reconstruct 8.28 -> 7.49 ms, d.outliers
// 4.32 -> 4.26 ms; bit-identical under the emulator and in the GPU parity suite. On.
#ifndef SPERR_REC_PREFIX
#define SPERR_REC_PREFIX 1
#endif

// counts[(c * maxp + n) * nblk + blk] = coefficients of unit blk significant before plane n
// (the caller zeroes the array: empty units write nothing)
__global__ void __launch_bounds__(kRecBlock, 2) k_rec_count(const DecChunk* jobs, unsigned* counts, int maxp,
                                                            unsigned nblk)
{
  __shared__ unsigned s_cnt[kMaxPlanes];
  __shared__ int s_max;
  __shared__ unsigned s_mask;
  const unsigned c = blockIdx.y, blk0 = blockIdx.x * kRecUnits;
  const DecChunk& d = jobs[c];
  if (d.skip || (unsigned long long)blk0 * kRecBlock >= d.n)
    return;
  int pu[kRecUnits];
  unsigned mine = 0;
#pragma unroll
  for (int u = 0; u < kRecUnits; u++) {
    const unsigned long long i = (unsigned long long)(blk0 + u) * kRecBlock + threadIdx.x;
    pu[u] = i < d.n ? rec_plane(gptr(d.pl)[i]) : -1;
  }
#pragma unroll
  for (int u = 0; u < kRecUnits; u++)
    mine |= pu[u] > 0 ? 1u << u : 0u;   // p == 0: significant, but never "before" any plane
  if (threadIdx.x == 0)
    s_mask = 0;
  __syncthreads();
  mine = __reduce_or_sync(0xffffffffu, mine);
  if ((threadIdx.x & 31) == 0 && mine)
    atomicOr(&s_mask, mine);
  __syncthreads();
  const unsigned mask = s_mask;
#pragma unroll
  for (int u = 0; u < kRecUnits; u++) {
    if (!((mask >> u) & 1u))
      continue;   // uniform over the block
    const int p = pu[u];
    if (threadIdx.x < kMaxPlanes)
      s_cnt[threadIdx.x] = 0;
    if (threadIdx.x == 0)
      s_max = -1;
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, p);
    if ((threadIdx.x & 31) == 0 && wmax >= 0)
      atomicMax(&s_max, wmax);
    __syncthreads();
    const int bmax = s_max;
    for (int n = 0; n < wmax; n++) {   // warp-uniform: nothing of this warp is significant before plane >= wmax
      const unsigned b = __ballot_sync(0xffffffffu, p > n);
      if ((threadIdx.x & 31) == 0 && b)
        atomicAdd(&s_cnt[n], unsigned(__popc(b)));
    }
    __syncthreads();
    if (int(threadIdx.x) < bmax)
      counts[((size_t)c * maxp + threadIdx.x) * nblk + blk0 + u] = s_cnt[threadIdx.x];
  }
}

// one block per (chunk, plane) row: exclusive scan over the blocks, in place
__global__ void k_rec_scan(const DecChunk* jobs, unsigned* counts, int maxp, unsigned nblk)
{
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry_s;
  const unsigned c = blockIdx.y, n = blockIdx.x;
  const DecChunk& d = jobs[c];
  if (d.skip || int(n) >= d.planes || d.ref_cnt[n] == 0)
    return;
  unsigned* row = counts + ((size_t)c * maxp + n) * nblk;
  const unsigned used = unsigned((d.n + kRecBlock - 1) / kRecBlock);
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < used; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < used ? row[i] : 0;
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o)
        inc += t;
    }
    if (lane == 31)
      wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned w = wsum[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o)
          w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    const unsigned excl = carry + inc - v + (warp ? wsum[warp - 1] : 0);
    if (i < used)
      row[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0)
      carry_s = carry + wsum[31];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kRecBlock, 2) k_rec_apply(const DecChunk* jobs, const ChunkDev* chunks,
                                                            const unsigned* counts, int maxp, unsigned nblk,
                                                            int mode, const double* tols, OutlierSink sink)
{
  __shared__ unsigned s_cnt[kMaxPlanes][32];   // per plane: significant-before-n count of every warp
  // per plane: refinement bits available and where the section starts (one global round trip per
  // block, in flight with the loads of the coefficients' states)
  __shared__ unsigned long long s_nref[kMaxPlanes], s_base[kMaxPlanes];
  __shared__ int s_max;
  __shared__ unsigned s_mask;
  const unsigned c = blockIdx.y, blk0 = blockIdx.x * kRecUnits;
  const DecChunk& d = jobs[c];
  const ChunkDev& ch = chunks[c];
  if (d.skip) {
    // no SPECK stream: every coefficient is zero (constant chunks never read their buffer)
    if (mode == 0 && !ch.is_const)
      for (int u = 0; u < kRecUnits; u++) {
        const unsigned long long i = (unsigned long long)(blk0 + u) * kRecBlock + threadIdx.x;
        if (i < ch.n)
          gptr(ch.coef)[i] = __dmul_rn(__dmul_rn(ch.q, 0.0), 1.0);
      }
    return;
  }
  if ((unsigned long long)blk0 * kRecBlock >= d.n)
    return;
  unsigned vu[kRecUnits];
#pragma unroll
  for (int u = 0; u < kRecUnits; u++) {
    const unsigned long long i = (unsigned long long)(blk0 + u) * kRecBlock + threadIdx.x;
    vu[u] = i < d.n ? gptr(d.pl)[i] : 0xFFu;
  }
  if (int(threadIdx.x) < d.planes && threadIdx.x < kMaxPlanes) {
    s_nref[threadIdx.x] = d.ref_cnt[threadIdx.x];
    s_base[threadIdx.x] = d.ref_base[threadIdx.x];
  }
  // units that need the full path: something has refinement bits (mode 0) / is an outlier (mode 1)
  unsigned full = 0;
#pragma unroll
  for (int u = 0; u < kRecUnits; u++)
    full |= rec_plane(vu[u]) >= (mode == 0 ? 1 : 0) ? 1u << u : 0u;
  if (threadIdx.x == 0)
    s_mask = 0;
  __syncthreads();
  full = __reduce_or_sync(0xffffffffu, full);
  if ((threadIdx.x & 31) == 0 && full)
    atomicOr(&s_mask, full);
  __syncthreads();
  const unsigned mask = s_mask;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
#pragma unroll
  for (int u = 0; u < kRecUnits; u++) {
    const unsigned blk = blk0 + u;
    const unsigned long long i = (unsigned long long)blk * kRecBlock + threadIdx.x;
    const unsigned v = vu[u];
    const int p = rec_plane(v);
    const bool in = i < d.n;
    const bool neg = p >= 0 && (v & 0x80u);
    if (!((mask >> u) & 1u)) {   // uniform over the block: all zero, or significant at plane 0 only
      if (mode == 0 && in)
        gptr(ch.coef)[i] = __dmul_rn(__dmul_rn(ch.q, p >= 0 ? 1.0 : 0.0), neg ? -1.0 : 1.0);
      continue;
    }
    if (threadIdx.x == 0)
      s_max = -1;
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, p);
    if (lane == 0 && wmax >= 0)
      atomicMax(&s_max, wmax);
    __syncthreads();
    const int bmax = s_max;
    for (int n = 0; n < bmax; n++) {
      const unsigned b = __ballot_sync(0xffffffffu, p > n);
      if (lane == 0)
        s_cnt[n][warp] = unsigned(__popc(b));
    }
    __syncthreads();
#if SPERR_REC_PREFIX
    for (int n = warp; n < bmax; n += 32) {   // warp-uniform
      const unsigned v = s_cnt[n][lane];
      unsigned inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
          inc += t;
      }
      const unsigned base = __shfl_sync(0xffffffffu, lane == 0 ? counts[((size_t)c * maxp + n) * nblk + blk] : 0u, 0);
      s_cnt[n][lane] = base + inc - v;
    }
    __syncthreads();
#endif
    unsigned long long mag = 0;
    if (p >= 0) {
      const unsigned long long thr = 1ull << p;
      mag = thr + thr - (thr >> 1) - 1ull;   // value given to a newly significant coefficient
    }
    for (int n = min(wmax, bmax) - 1; n >= 0; n--) {   // warp-uniform: planes below this warp's largest
      const unsigned long long nref = s_nref[n];
      if (nref == 0)
        continue;   // plane never refined (not reached, or the stream ended before its section)
      const bool mine = p > n;
      const unsigned b = __ballot_sync(0xffffffffu, mine);
      if (b == 0)
        continue;
#if SPERR_REC_PREFIX
      const unsigned long long first = s_cnt[n][warp];
#else
      const unsigned before = __reduce_add_sync(0xffffffffu, lane < warp ? s_cnt[n][lane] : 0u);
#endif
      if (mine) {
#if SPERR_REC_PREFIX
        const unsigned long long rank = first + __popc(b & lt);
#else
        const unsigned long long rank =
            (unsigned long long)counts[((size_t)c * maxp + n) * nblk + blk] + before + __popc(b & lt);
#endif
        if (rank < nref) {
          const unsigned long long bp = s_base[n] + rank;
          const unsigned bit = (gptr(d.bits)[bp >> 5] >> (bp & 31)) & 1u;
          if (n >= 1) {
            const unsigned long long half = 1ull << (n - 1);
            mag = bit ? mag + half : mag - half;
          }
          else
            mag += bit;
        }
      }
    }
    if (mode == 0) {
      if (in)
        gptr(ch.coef)[i] = __dmul_rn(__dmul_rn(ch.q, __ull2double_rn(mag)), neg ? -1.0 : 1.0);
    }
    else {
      // outlier correctors: a sorted list for the consumers plus one flag bit per value
      const bool has = in && p >= 0 && mag != 0;
      const unsigned b = __ballot_sync(0xffffffffu, has);
      if (lane == 0 && in)
        gptr(ch.obits)[i >> 5] = b;
      if (has) {
        double e = mag == 1 ? 1.1 : __dsub_rn(__ull2double_rn(mag), 0.25);
        e = __dmul_rn(e, __dmul_rn(tols[c], neg ? -1.0 : 1.0));
        outlier_append(sink, c, i, e);
      }
    }
    // s_cnt / s_max are rewritten by the next unit only after its first barrier
  }
}

// coef[pos] += corrector for the chunks that are not handled by the fused inverse transform
__global__ void k_apply_correctors(const ChunkDev* chunks, const unsigned long long* key,
                                   const double* val, unsigned long long n)
{
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n)
    return;
  const unsigned long long k = key[i];
  const ChunkDev& ch = chunks[unsigned(k >> 32)];
  if (ch.fused)
    return;
  const unsigned long long pos = k & 0xffffffffull;
  gptr(ch.coef)[pos] = __dadd_rn(gptr(ch.coef)[pos], val[i]);
}

void launch_apply_correctors(const ChunkDev* d_chunks, const unsigned long long* d_key,
                             const double* d_val, unsigned long long n, cudaStream_t st)
{
  if (n)
    LAUNCH(k_apply_correctors, dim3(unsigned((n + 255) / 256)), dim3(256), 0, st, d_chunks, d_key, d_val, n);
}

void speck_reconstruct(DecWork& w, const ChunkDev* d_chunks, int mode, const double* d_tols,
                       int first, int count, const OutlierSink& sink, cudaStream_t st)
{
  const int nj = count;
  if (nj == 0)
    return;
  const DecChunk* dj = w.dchunks.as<DecChunk>() + first;
  // mode 0 must also zero-fill chunks without a stream, whose extent the jobs do not know
  size_t max_n = w.max_n;
  if (mode == 0)
    max_n = std::max(max_n, w.fill_n);
  if (max_n == 0)
    return;
  const unsigned nblk = unsigned((max_n + kRecBlock - 1) / kRecBlock);
  const int maxp = std::max(1, w.max_planes);
  const size_t ncounts = (size_t)nj * maxp * nblk;
  w.counts.reserve(ncounts * 4);
  rt::dset(w.counts.p, 0, ncounts * 4, st);
  unsigned* cnt = w.counts.as<unsigned>();
  if (w.max_n) {
    LAUNCH(k_rec_count, dim3((nblk + kRecUnits - 1) / kRecUnits, nj), dim3(kRecBlock), 0, st, dj, cnt, maxp, nblk);
    LAUNCH(k_rec_scan, dim3(maxp, nj), dim3(1024), 0, st, dj, cnt, maxp, nblk);
  }
  LAUNCH(k_rec_apply, dim3((nblk + kRecUnits - 1) / kRecUnits, nj), dim3(kRecBlock), 0, st, dj, d_chunks, cnt,
         maxp, nblk, mode, d_tols, sink);
}

}  // namespace sperr_b200
