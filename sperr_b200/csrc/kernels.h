// Declarations of the kernel launchers (one translation unit per stage).
#pragma once

#include <algorithm>
#include <vector>

#include "dev.h"

namespace sperr_b200 {

// A strided 3D array in device memory (source volume on compress, destination on decompress).
struct SrcVol {
  const void* ptr;
  int is_float;
  unsigned long long vx, vy;   // row length and number of rows per plane
  unsigned long long vz = 0;   // number of planes (0: unknown; only the TMA path of the forward transform needs it)
};

struct CdfC {
  double ALPHA, BETA, GAMMA, DELTA, EPSILON, INV_EPSILON;
  int fma;   // arithmetic flavour of the lifting steps (see lift_add below)
};

// Arithmetic flavour of the lifting steps. STRICT (default): every multiply and add rounded
// separately, the reference built with -ffp-contract=off. FMA: the contraction pattern of the
// reference's stock x86 build (g++ -O3 -mfma, -ffp-contract=fast; SURVEY.md appendix A.3, pinned
// against that build in tests/test_fma_flavour.py): x + C*s and x - C*s become one fma, the scaled
// steps keep their separate multiply. The flavour is a template parameter of the hot kernels, so
// the STRICT code is unchanged by the existence of the other one.
int& fma_flavour();   // process-wide switch read by cdf_constants(): 0 STRICT, 1 FMA
template <bool FMA>
__device__ __forceinline__ double lift_add(double x, double C, double s)   // x + C * s
{
  if (FMA)
    return __fma_rn(C, s, x);
  return __dadd_rn(x, __dmul_rn(C, s));
}
template <bool FMA>
__device__ __forceinline__ double lift_sub(double x, double C, double s)   // x - C * s
{
  if (FMA)
    return __fma_rn(-C, s, x);
  return __dsub_rn(x, __dmul_rn(C, s));
}
template <bool FMA>
__device__ __forceinline__ double lift_scale_fwd(const CdfC& k, double e, double s)   // EPSILON * (e + DELTA * s)
{
  if (FMA)
    return __dmul_rn(k.EPSILON, __fma_rn(k.DELTA, s, e));
  return __dmul_rn(k.EPSILON, __dadd_rn(e, __dmul_rn(k.DELTA, s)));
}
template <bool FMA>
__device__ __forceinline__ double lift_scale_inv(const CdfC& k, double e, double s)   // e * INV_EPSILON - DELTA * s
{
  if (FMA)
    return __fma_rn(e, k.INV_EPSILON, -__dmul_rn(k.DELTA, s));
  return __dsub_rn(__dmul_rn(e, k.INV_EPSILON), __dmul_rn(k.DELTA, s));
}

// Unordered append list of PWE outliers found by a batch: key = chunk << 32 | position in chunk.
struct OutlierSink {
  unsigned long long* total;   // entries appended so far (may exceed cap: the caller retries)
  unsigned* per_chunk;         // entries per chunk
  unsigned long long* key;
  double* err;
  unsigned long long cap;
};
__device__ __forceinline__ void outlier_append(const OutlierSink& s, unsigned chunk, unsigned long long pos,
                                               double err)
{
  atomicAdd(&s.per_chunk[chunk], 1u);
  const unsigned long long slot = atomicAdd(s.total, 1ull);
  if (slot < s.cap) {
    s.key[slot] = ((unsigned long long)chunk << 32) | pos;
    s.err[slot] = err;
  }
}

// Outlier correctors of a batch being decoded, sorted by key (chunk << 32 | position); chunk c's
// entries are [off[c], off[c + 1]). key == nullptr: no chunk has correctors.
struct CorrectorList {
  const unsigned long long* key;
  const double* val;
  const unsigned long long* off;
};
__device__ __forceinline__ double corrector_lookup(const CorrectorList& l, unsigned chunk, unsigned long long pos)
{
  const unsigned long long want = ((unsigned long long)chunk << 32) | pos;
  unsigned long long lo = l.off[chunk], hi = l.off[chunk + 1];
  while (lo < hi) {
    const unsigned long long mid = (lo + hi) >> 1;
    const unsigned long long k = l.key[mid];
    if (k == want)
      return l.val[mid];
    if (k < want)
      lo = mid + 1;
    else
      hi = mid;
  }
  return 0.0;
}

// ---- dwt_fused.cu: one HBM round trip per level, dyadic chunks only ----
size_t fused_scratch_elems(uint32_t nx, uint32_t ny, uint32_t nz, long long off[8]);
void launch_dwt_fused_forward(const SrcVol& src, const ChunkDev* d_chunks, const int* d_ids, int nids,
                              uint32_t nx, uint32_t ny, uint32_t nz, cudaStream_t st);
void launch_dwt_fused_inverse(const SrcVol& vol, int mode, const ChunkDev* d_chunks, const int* d_ids,
                              int nids, uint32_t nx, uint32_t ny, uint32_t nz, double tol,
                              const OutlierSink& sink, const CorrectorList& cor, cudaStream_t st);
void launch_level_gather(const ChunkDev* d_chunks, int nchunks, uint32_t nx, uint32_t ny, uint32_t nz,
                         int h, void* dst, int is_float, size_t vx, size_t vy, cudaStream_t st);
CdfC cdf_constants();

// ---- transform.cu ----
void launch_stats(const SrcVol& src, ChunkDev* d_chunks, int nchunks, double* d_stride_mean,
                  int max_strides, const unsigned* d_nstrides, unsigned* d_not_const,
                  bool want_minmax, cudaStream_t st);
void launch_gather(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                   cudaStream_t st);
void launch_dwt(bool inverse, const ChunkDev* d_chunks, const int* d_ids, int nids, uint32_t nx,
                uint32_t ny, uint32_t nz, bool is_2d, cudaStream_t st);
void launch_absmax(ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st);
void launch_qdecide(ChunkDev* d_chunks, int nchunks, cudaStream_t st);
void launch_quantize(const ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st);
void launch_inv_quantize(const ChunkDev* d_chunks, int nchunks, size_t max_n, cudaStream_t st);
void launch_mse(const ChunkDev* d_chunks, const int* d_ids, const double* d_qs, double* d_partial,
                int max_strides, double* d_mse, int nslots, cudaStream_t st);
void launch_scatter_out(const SrcVol& dst, const ChunkDev* d_chunks, int nchunks, size_t max_n,
                        cudaStream_t st);

// ---- primitives.cu ----
// out[i] = sum_{j<i} in[j] for i in [0, n]  (n+1 outputs). `tmp` needs scan_tmp_bytes(n).
size_t scan_tmp_bytes(size_t n);
void exclusive_scan_u32(const unsigned* d_in, unsigned long long* d_out, size_t n, void* d_tmp,
                        cudaStream_t st);
// Sorts (key, value) pairs by the low `bits` bits of the key.
size_t sort_tmp_bytes(size_t n);
void sort_pairs_u64(const unsigned long long* kin, unsigned long long* kout,
                    const unsigned long long* vin, unsigned long long* vout, size_t n, int bits,
                    void* d_tmp, size_t tmp_bytes, cudaStream_t st);

}  // namespace sperr_b200
