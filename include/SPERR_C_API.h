/*
 * SPERR's C API as implemented by libsperr_b200.so: the same six symbols, argument lists and return
 * codes as the reference's header (/root/reference/include/SPERR_C_API.h:53-156, behaviour in
 * src/SPERR_C_API.cpp:11-281), so C / Fortran / Python callers and filters such as H5Z-SPERR build
 * against this file and link -lsperr_b200 instead of -lSPERR without source changes.
 *
 * All buffers are HOST memory. Output buffers: `*dst` must be NULL on entry (else the call returns
 * 1); on success it points to a malloc()ed buffer the caller free()s. Modes: 1 fixed rate in bits
 * per value, 2 target PSNR, 3 point-wise error tolerance; `quality` <= 0 or another mode returns 2;
 * every other failure (bad stream, no CUDA device, ...) returns -1. `nthreads` (OpenMP threads in
 * the reference) is accepted and ignored: chunks are coded on the GPU(s).
 * GPU-only additions (device pointers, batched 2D slices, chunk ranges) are in sperr_b200.h.
 */
#ifndef SPERR_C_API_H
#define SPERR_C_API_H

#ifndef USE_VANILLA_CONFIG
#include "SperrConfig.h"
#endif

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
namespace C_API {
extern "C" {
#endif

/* One 2D slice (x fastest) -> stream; out_inc_header = 1 prepends the 10-byte slice header. */
int sperr_comp_2d(const void* src, int is_float, size_t dimx, size_t dimy, int mode, double quality,
                  int out_inc_header, void** dst, size_t* dst_len);

/* Headerless slice stream -> dimx * dimy floats (output_float = 1) or doubles. */
int sperr_decomp_2d(const void* src, size_t src_len, int output_float, size_t dimx, size_t dimy,
                    void** dst);

/* Dimensions and input precision out of a 3D container or a 2D slice header (dimz = 1). */
void sperr_parse_header(const void* src, size_t* dimx, size_t* dimy, size_t* dimz, int* is_float);

/* 3D volume (x fastest, z slowest) -> container; chunk_* are preferred chunk extents. */
int sperr_comp_3d(const void* src, int is_float, size_t dimx, size_t dimy, size_t dimz, size_t chunk_x,
                  size_t chunk_y, size_t chunk_z, int mode, double quality, size_t nthreads, void** dst,
                  size_t* dst_len);

/* Container -> volume of floats (output_float = 1) or doubles; dimensions are returned. */
int sperr_decomp_3d(const void* src, size_t src_len, int output_float, size_t nthreads, size_t* dimx,
                    size_t* dimy, size_t* dimz, void** dst);

/* Keeps about pct % (1..100) of every chunk of a container (at least 64 bytes per chunk); the
 * result is itself a decodable container. src_len may be shorter than the full container as long
 * as it covers what is kept. */
int sperr_trunc_3d(const void* src, size_t src_len, unsigned pct, void** dst, size_t* dst_len);

#ifdef __cplusplus
} /* extern "C" */
} /* namespace C_API */
#endif

#endif
