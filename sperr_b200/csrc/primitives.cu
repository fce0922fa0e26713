// Small device-wide primitives: exclusive scan and key/value radix sort.
#include "kernels.h"

#ifndef SPERR_EMUL
#include <cub/cub.cuh>
#endif

namespace sperr_b200 {

constexpr int kScanBlock = 1024;

// Block-level inclusive scan of 64-bit values (one per thread, 1024 threads).
__device__ __forceinline__ unsigned long long block_incl_scan(unsigned long long v,
                                                              unsigned long long* total)
{
  __shared__ unsigned long long wsum[32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o)
      v += t;
  }
  if (lane == 31)
    wsum[warp] = v;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = wsum[lane];
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o)
        w += t;
    }
    wsum[lane] = w;
  }
  __syncthreads();
  if (warp > 0)
    v += wsum[warp - 1];
  *total = wsum[31];
  __syncthreads();
  return v;
}

__global__ void k_scan_blocks(const unsigned* in, unsigned long long* out, size_t n,
                              unsigned long long* block_tot)
{
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  const unsigned long long v = i < n ? in[i] : 0;
  unsigned long long tot;
  const unsigned long long inc = block_incl_scan(v, &tot);
  if (i < n)
    out[i] = inc - v;  // exclusive, block-local
  if (threadIdx.x == 0)
    block_tot[blockIdx.x] = tot;
}

// Single block: exclusive scan of the block totals, in place; element [nb] receives the sum.
__global__ void k_scan_totals(unsigned long long* block_tot, size_t nb)
{
  unsigned long long carry = 0;
  for (size_t base = 0; base < nb; base += kScanBlock) {
    const size_t i = base + threadIdx.x;
    const unsigned long long v = i < nb ? block_tot[i] : 0;
    unsigned long long tot;
    const unsigned long long inc = block_incl_scan(v, &tot);
    if (i < nb)
      block_tot[i] = carry + inc - v;
    carry += tot;
  }
  if (threadIdx.x == 0)
    block_tot[nb] = carry;
}

__global__ void k_scan_add(unsigned long long* out, size_t n, const unsigned long long* block_tot,
                           size_t nb)
{
  const size_t i = (size_t)blockIdx.x * kScanBlock + threadIdx.x;
  if (i < n)
    out[i] += block_tot[blockIdx.x];
  if (i == 0)
    out[n] = block_tot[nb];
}

size_t scan_tmp_bytes(size_t n)
{
  return ((n + kScanBlock - 1) / kScanBlock + 2) * sizeof(unsigned long long);
}

void exclusive_scan_u32(const unsigned* d_in, unsigned long long* d_out, size_t n, void* d_tmp,
                        cudaStream_t st)
{
  unsigned long long* tot = reinterpret_cast<unsigned long long*>(d_tmp);
  const size_t nb = (n + kScanBlock - 1) / kScanBlock;
  if (n == 0) {
    rt::dset(d_out, 0, sizeof(unsigned long long), st);
    return;
  }
  LAUNCH(k_scan_blocks, dim3(unsigned(nb)), dim3(kScanBlock), 0, st, d_in, d_out, n, tot);
  LAUNCH(k_scan_totals, dim3(1), dim3(kScanBlock), 0, st, tot, nb);
  LAUNCH(k_scan_add, dim3(unsigned(nb)), dim3(kScanBlock), 0, st, d_out, n, tot, nb);
}

#ifndef SPERR_EMUL
size_t sort_tmp_bytes(size_t n)
{
  size_t bytes = 0;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, (const unsigned long long*)nullptr,
                                  (unsigned long long*)nullptr, n > 0 ? n : 1, 0, 64);
  return bytes;
}

void sort_pairs_u64(const unsigned long long* kin, unsigned long long* kout,
                    const unsigned long long* vin, unsigned long long* vout, size_t n, int bits,
                    void* d_tmp, size_t tmp_bytes, cudaStream_t st)
{
  if (n == 0)
    return;
  RT_CHECK(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, kin, kout, vin, vout, n, 0, bits, st));
}
#else
size_t sort_tmp_bytes(size_t) { return 16; }

void sort_pairs_u64(const unsigned long long* kin, unsigned long long* kout,
                    const unsigned long long* vin, unsigned long long* vout, size_t n, int bits,
                    void*, size_t, cudaStream_t)
{
  std::vector<size_t> idx(n);
  for (size_t i = 0; i < n; i++)
    idx[i] = i;
  const unsigned long long mask = bits >= 64 ? ~0ull : ((1ull << bits) - 1);
  std::stable_sort(idx.begin(), idx.end(),
                   [&](size_t a, size_t b) { return (kin[a] & mask) < (kin[b] & mask); });
  for (size_t i = 0; i < n; i++) {
    kout[i] = kin[idx[i]];
    vout[i] = vin[idx[i]];
  }
}
#endif

}  // namespace sperr_b200
