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
};

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
