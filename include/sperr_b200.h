/*
 * sperr_b200 -- C ABI of the B200-native SPERR hot path (libsperr_b200.so).
 *
 * Section 1 are drop-in replacements of the reference's C API: same names, argument meaning,
 * ownership rules and return codes as /root/reference/include/SPERR_C_API.h:53-156, so a program
 * (or an FFI binding: H5Z-SPERR, the Fortran wrapper, ctypes) links against this library instead
 * of libSPERR without source changes. Buffers are HOST memory unless a name ends in _dev.
 *
 * Section 2 are additive GPU-side entry points (device-resident data, chunk-range sharding for
 * one-process-per-GPU jobs); section 3 are stage-level hooks used by the parity tests.
 *
 * All functions are synchronous and thread-compatible (one call at a time per process is the
 * tested configuration). The library has no CPU execution path: without a CUDA device every
 * compute entry point returns -1.
 */
#ifndef SPERR_B200_H
#define SPERR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------ */
/* 1. Reference-compatible API                                                                 */
/* ------------------------------------------------------------------------------------------ */

/* Replaces C_API::sperr_comp_3d (SPERR_C_API.h:84-111, src/SPERR_C_API.cpp:156-216).
 * mode: 1 = fixed bit-per-pixel, 2 = fixed PSNR, 3 = fixed point-wise error.
 * `nthreads` is accepted for signature compatibility and ignored (the chunk loop runs on the GPU).
 * Returns 0 ok; 1 *dst not NULL; 2 bad parameter; -1 other error. *dst is malloc'd. */
int sperr_comp_3d(const void* src, int is_float, size_t dimx, size_t dimy, size_t dimz,
                  size_t chunk_x, size_t chunk_y, size_t chunk_z, int mode, double quality,
                  size_t nthreads, void** dst, size_t* dst_len);

/* Replaces C_API::sperr_decomp_3d (SPERR_C_API.h:113-131, src/SPERR_C_API.cpp:218-258). */
int sperr_decomp_3d(const void* src, size_t src_len, int output_float, size_t nthreads, size_t* dimx,
                    size_t* dimy, size_t* dimz, void** dst);

/* Replaces C_API::sperr_parse_header (SPERR_C_API.h:74-82, src/SPERR_C_API.cpp:136-154). */
void sperr_parse_header(const void* src, size_t* dimx, size_t* dimy, size_t* dimz, int* is_float);

/* Replaces C_API::sperr_trunc_3d (SPERR_C_API.h:133-156, src/SPERR_C_API.cpp:260-281): keeps the
 * first pct % of every chunk's stream (progressive access, SPERR3D_Stream_Tools.cpp:131-226) and
 * marks the container as a portion. pct == 0 or >= 100 copies the stream. Host bytes only: needs no
 * GPU. Returns 0 ok; 1 *dst not NULL; -1 malformed / short input. *dst is malloc'd. */
int sperr_trunc_3d(const void* src, size_t src_len, unsigned pct, void** dst, size_t* dst_len);

/* Replaces C_API::sperr_comp_2d (SPERR_C_API.h:33-60, src/SPERR_C_API.cpp:11-95): one 2D slice,
 * x fastest. out_inc_header != 0 prepends the 10-byte 2D header {version, flags, u32 dimx, u32 dimy}.
 * Returns 0 ok; 1 *dst not NULL; 2 bad parameter; -1 other error. *dst is malloc'd. */
int sperr_comp_2d(const void* src, int is_float, size_t dimx, size_t dimy, int mode, double quality,
                  int out_inc_header, void** dst, size_t* dst_len);

/* Replaces C_API::sperr_decomp_2d (SPERR_C_API.h:62-72, src/SPERR_C_API.cpp:97-134): `src` is the
 * stream WITHOUT the 10-byte header. */
int sperr_decomp_2d(const void* src, size_t src_len, int output_float, size_t dimx, size_t dimy,
                    void** dst);

/* ------------------------------------------------------------------------------------------ */
/* 2. GPU-side extensions                                                                      */
/* ------------------------------------------------------------------------------------------ */

/* Same as sperr_comp_3d but `src` is a DEVICE pointer to the whole volume and the output container
 * is written to a malloc'd HOST buffer. */
int sperr_b200_comp_3d_dev(const void* d_src, int is_float, size_t dimx, size_t dimy, size_t dimz,
                           size_t chunk_x, size_t chunk_y, size_t chunk_z, int mode, double quality,
                           void** dst, size_t* dst_len);

/* Decompresses a reference-layout 3D container. `h_src` is the container in HOST memory (headers are
 * parsed on the host); `d_src` is the same bytes in DEVICE memory, or NULL to have them uploaded.
 * The decoded volume (float when output_float != 0, else double) is written to the DEVICE buffer
 * d_dst, which must hold dimx*dimy*dimz values (see sperr_parse_header). */
int sperr_b200_decomp_3d_dev(const void* h_src, const void* d_src, size_t src_len, int output_float,
                             size_t* dimx, size_t* dimy, size_t* dimz, void* d_dst);

/* Multi-resolution decoding (what sperr::SPERR3D_OMP_D::decompress(p, true) offers,
 * include/SPERR3D_OMP_D.h:24-29, src/SPERR3D_OMP_D.cpp:70-126): besides the full volume (*dst, as
 * sperr_decomp_3d) the coarsened volumes the inverse wavelet transform passes through, coarsest
 * first. They exist when the volume is a whole number of chunks and the chunk shape has a dyadic
 * transform (sperr::coarsened_resolutions, src/sperr_helper.cpp:70-123); otherwise *nlevels = 0.
 * level_dims receives 3 extents per level (room for 8 levels), level_data one malloc'd buffer per
 * level (room for 8 pointers; float or double like *dst; the caller frees them). */
int sperr_b200_decomp_3d_multires(const void* src, size_t src_len, int output_float, size_t* dimx,
                                  size_t* dimy, size_t* dimz, void** dst, size_t* nlevels,
                                  size_t* level_dims, void** level_data);

/* ---- 2a. batched 2D slices ----
 * sperr_comp_2d codes one slice per call; a GPU wants many slices in flight. These entry points
 * code `nslices` independent slices of dimx x dimy (contiguous, slice s at src + s*dimx*dimy) in one
 * call: stream s is byte-identical to what sperr_comp_2d returns for slice s. The streams are
 * written back to back into one malloc'd buffer *dst; lens[s] receives the length of stream s.
 * `_dev`: src is a DEVICE pointer. */
int sperr_b200_comp_2d_batch(const void* src, int is_float, size_t dimx, size_t dimy, size_t nslices,
                             int mode, double quality, int out_inc_header, void** dst,
                             size_t* lens);
int sperr_b200_comp_2d_batch_dev(const void* d_src, int is_float, size_t dimx, size_t dimy,
                                 size_t nslices, int mode, double quality, int out_inc_header,
                                 void** dst, size_t* lens);
/* Decodes `nslices` headerless slice streams (back to back in src, lens[s] each) into one malloc'd
 * buffer *dst of nslices*dimx*dimy values (`_dev`: into the caller's DEVICE buffer d_dst). */
int sperr_b200_decomp_2d_batch(const void* src, const size_t* lens, size_t nslices, int output_float,
                               size_t dimx, size_t dimy, void** dst);
int sperr_b200_decomp_2d_batch_dev(const void* src, const size_t* lens, size_t nslices,
                                   int output_float, size_t dimx, size_t dimy, void* d_dst);

/* ---- 2b. chunk-range sharding (one process per GPU) ----
 * SPERR's chunks are coded independently (SPERR3D_OMP_C.cpp:94-130, SPERR3D_OMP_D.cpp:94-130): rank r
 * codes chunks [chunk_begin, chunk_end) of chunk_volume's order (sperr_helper.cpp:542-592) out of
 * the bounding box of those chunks, which it holds in device memory (x fastest, row length
 * box_extent[0], box_extent[1] rows per plane). vol / chunk are the whole volume's extents and the
 * preferred chunk extents, exactly as given to sperr_comp_3d. */

/* Number of chunks of the volume. */
size_t sperr_b200_num_chunks(const size_t vol[3], const size_t chunk[3]);

/* Bounding box of chunks [begin, end). Returns 0, or -1 for a bad range. */
int sperr_b200_chunk_box(const size_t vol[3], const size_t chunk[3], size_t begin, size_t end,
                         size_t origin[3], size_t extent[3]);

/* Partition of the chunks over `world` ranks: rank r owns [begins[r], begins[r + 1]) (begins has
 * world + 1 entries). Every range is exactly a box of chunks, so a rank's bounding box holds no
 * chunk of another rank (even split of chunks, else of whole chunk rows, else of whole z-slabs).
 * Returns 0, or -1 when there are fewer chunks than ranks or no such split exists. */
int sperr_b200_shard_ranges(const size_t vol[3], const size_t chunk[3], size_t world, size_t* begins);

/* Compresses the chunks of the range. *d_streams receives a library-owned DEVICE buffer (valid
 * until the next call into the library) holding their streams back to back, lens[i] the length of
 * chunk chunk_begin + i. Returns 0 ok; 2 bad parameter; -1 other error. */
int sperr_b200_comp_3d_range_dev(const void* d_box, int is_float, const size_t vol[3],
                                 const size_t chunk[3], const size_t box_origin[3],
                                 const size_t box_extent[3], size_t chunk_begin, size_t chunk_end,
                                 int mode, double quality, const void** d_streams, size_t* streams_len,
                                 uint32_t* lens);

/* Synchronous copy helper for callers that hold raw pointers: kind 0 device -> device, 1 host ->
 * device, 2 device -> host. */
int sperr_b200_memcpy_dev(void* dst, const void* src, size_t n, int kind);

/* Writes the reference container header (SPERR3D_OMP_C::m_generate_header, SPERR3D_OMP_C.cpp:
 * 163-234) for `nchunks` chunk lengths into out[0, cap) and returns its length (call with out = NULL
 * to size the buffer). The container is this header followed by all chunk streams in order. */
size_t sperr_b200_container_header(const size_t vol[3], const size_t chunk[3], int is_float,
                                   const uint32_t* lens, size_t nchunks, void* out, size_t cap);

/* Parses a container header (SPERR3D_Stream_Tools::get_stream_header, SPERR3D_Stream_Tools.cpp:
 * 46-105). lens may be NULL; otherwise it receives *nchunks lengths (cap entries available). */
int sperr_b200_parse_container(const void* src, size_t len, size_t vol[3], size_t chunk[3],
                               int* is_float, size_t* header_len, uint32_t* lens, size_t cap,
                               size_t* nchunks);

/* Decodes the chunks of the range from their streams (back to back, lens[i] each) into the caller's
 * DEVICE box (float when output_float != 0, else double). h_streams: the streams in HOST memory
 * (chunk headers are parsed on the host), or NULL when only the device copy exists (the library then
 * fetches the 26 + 9 header bytes of every chunk itself); d_streams: the same bytes in DEVICE memory,
 * or NULL to have them uploaded. At least one of the two must be given. */
int sperr_b200_decomp_3d_range_dev(const void* h_streams, const void* d_streams, size_t streams_len,
                                   const uint32_t* lens,
                                   const size_t vol[3], const size_t chunk[3],
                                   const size_t box_origin[3], const size_t box_extent[3],
                                   size_t chunk_begin, size_t chunk_end, int output_float,
                                   void* d_box_out);

/* Arithmetic flavour of the CDF 9/7 lifting steps, process-wide, for all later calls. 0 (default):
 * every multiply and add rounded separately -- streams and decoded values equal to the reference
 * built with -ffp-contract=off (what the north star calls the reference's exact operation order
 * with -fmad=false). 1: the steps contracted the way the reference's stock x86 build contracts them
 * (g++ -O3 -mfma with GCC's default -ffp-contract=fast: x +- C*s as one fma), so that streams and
 * decoded values equal THAT library's (SURVEY.md section 0 finding 1: fixed-rate streams and the
 * last bit of decoded values differ between the two builds). Also settable with the environment
 * variable SPERR_B200_FMA=1 before the first call. */
void sperr_b200_set_fma_flavour(int on);

/* Stage profiler: when enabled every stage of the pipelines is bracketed by CUDA events on the
 * launching stream. prof_dump writes a JSON object {"stage": {"ms": total, "n": ranges}, ...} into
 * buf (NUL-terminated, truncated to cap) and returns the full length. Enabling clears the totals. */
void sperr_b200_prof_enable(int on);
size_t sperr_b200_prof_dump(char* buf, size_t cap);

/* Number of kernels of this library launched by the calling process so far (library kernels such
 * as cub's radix sort are not counted). */
unsigned long long sperr_b200_launch_count(void);

/* ------------------------------------------------------------------------------------------ */
/* 3. Stage-level hooks (parity tests)                                                         */
/* ------------------------------------------------------------------------------------------ */

int sperr_b200_stage_condition(const void* src, int is_float, size_t nx, size_t ny, size_t nz,
                               double* out_vals, double* out_mean, int* out_is_const);
int sperr_b200_stage_dwt(double* buf, size_t nx, size_t ny, size_t nz, int inverse, int is_2d);
/* Same transform through the fused one-round-trip-per-level kernels (dyadic shapes only; -2 else). */
int sperr_b200_stage_dwt_fused(double* buf, size_t nx, size_t ny, size_t nz, int inverse);
int sperr_b200_stage_quantize(const double* vals, size_t nx, size_t ny, size_t nz, double q,
                              uint64_t* mags, uint8_t* signs, int* wide);
int sperr_b200_stage_speck3d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                                    size_t nz, size_t budget_bits, uint8_t* out, size_t cap,
                                    size_t* out_len);

/* SPECK2D_INT_ENC on one slice (src/SPECK2D_INT*.cpp). */
int sperr_b200_stage_speck2d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                                    size_t budget_bits, uint8_t* out, size_t cap, size_t* out_len);

int sperr_b200_stage_outlier_encode(const uint64_t* pos, const double* err, size_t n_out,
                                    size_t total_len, double tol, uint8_t* out, size_t cap,
                                    size_t* out_len);

#ifdef __cplusplus
}
#endif
#endif
