/*
 * TEST INFRASTRUCTURE ONLY -- the parity oracle. Nothing in the product path (sperr_b200/,
 * include/) may include, link or call this. Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py use it, and only as the checker.
 *
 * A plain-C restatement of the SPERR v0.8.5 3D/2D compression hot path (STRICT arithmetic:
 * every fp64 multiply and add individually rounded, i.e. the reference compiled with
 * -ffp-contract=off). Each function in sperr_oracle.c cites the reference file:line it follows.
 *
 * Parity status: PINNED. tests/test_oracle_vs_ref.py checks every stage and the end-to-end
 * streams / decoded values of this restatement for bit equality against oracle/_ref/libsperr_ref.so,
 * which is the unmodified reference compiled by oracle/Makefile, and tests/test_oracle_kat.py
 * checks the known-answer tables of the reference's own unit tests.
 */
#ifndef SPERR_ORACLE_H
#define SPERR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- geometry (src/sperr_helper.cpp) ---- */
size_t so_num_of_xforms(size_t len);
size_t so_num_of_partitions(size_t len);
int so_can_use_dyadic(size_t nx, size_t ny, size_t nz); /* levels, or -1 */
void so_calc_approx_detail_len(size_t len, size_t lev, size_t out2[2]);
size_t so_chunk_volume(size_t vx, size_t vy, size_t vz, size_t cx, size_t cy, size_t cz,
                       size_t* out6, size_t cap);

/* ---- conditioner (src/Conditioner.cpp) ---- */
int so_condition(double* buf, size_t n, uint8_t header17[17]); /* returns 1 if constant */
void so_inverse_condition(double* buf, size_t n, const uint8_t header17[17]);

/* ---- CDF 9/7 (src/CDF97.cpp) ---- */
void so_dwt3d(double* buf, size_t nx, size_t ny, size_t nz);
void so_idwt3d(double* buf, size_t nx, size_t ny, size_t nz);
void so_dwt2d(double* buf, size_t nx, size_t ny);
void so_idwt2d(double* buf, size_t nx, size_t ny);

/* ---- mid-tread quantiser (src/SPECK_FLT.cpp:311-399) ---- */
/* returns 0 ok, -1 FE_INVALID; width = 1,2,4,8 */
int so_quantize(const double* vals, size_t n, double q, uint64_t* mag, uint8_t* sign, int* width);
void so_inv_quantize(const uint64_t* mag, const uint8_t* sign, size_t n, double q, double* vals);
double so_estimate_mse_midtread(const double* vals, size_t n, double q);

/* ---- integer SPECK (src/SPECK_INT.cpp, src/SPECK3D_INT*.cpp, src/SPECK1D_INT*.cpp) ---- */
/* budget_bits == 0 means unlimited. Returns stream bytes (written only if <= cap). */
size_t so_speck3d_encode(const uint64_t* mag, const uint8_t* sign, size_t nx, size_t ny, size_t nz,
                         size_t budget_bits, uint8_t* out, size_t cap);
void so_speck3d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, size_t nz,
                       uint64_t* mag, uint8_t* sign);
size_t so_speck1d_encode(const uint64_t* mag, const uint8_t* sign, size_t n, uint8_t* out,
                         size_t cap);
void so_speck1d_decode(const uint8_t* stream, size_t len, size_t n, uint64_t* mag, uint8_t* sign);
size_t so_speck2d_encode(const uint64_t* mag, const uint8_t* sign, size_t nx, size_t ny,
                         size_t budget_bits, uint8_t* out, size_t cap);
void so_speck2d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, uint64_t* mag,
                       uint8_t* sign);

/* ---- outlier coder (src/Outlier_Coder.cpp) ---- */
size_t so_outlier_encode(const uint64_t* pos, const double* err, size_t n_out, size_t total_len,
                         double tol, uint8_t* out, size_t cap);
size_t so_outlier_decode(const uint8_t* stream, size_t len, size_t total_len, double tol,
                         uint64_t* pos, double* err, size_t cap);

/* ---- one chunk (src/SPECK_FLT.cpp:401-606); mode 1 = BPP, 2 = PSNR, 3 = PWE ---- */
/* `vals` (fp64, already widened) is consumed. is_2d selects dwt2d + SPECK2D. */
/* Returns bytes written to a malloc'd *out, or 0 on error. */
size_t so_chunk_compress(double* vals, size_t nx, size_t ny, size_t nz, int mode, double quality,
                         int is_2d, uint8_t** out);
int so_chunk_decompress(const uint8_t* stream, size_t len, size_t nx, size_t ny, size_t nz,
                        int is_2d, double* out);

/* ---- whole volume: same contracts as the reference C API (include/SPERR_C_API.h:53-156) ---- */
int so_comp_3d(const void* src, int is_float, size_t dimx, size_t dimy, size_t dimz, size_t chunk_x,
               size_t chunk_y, size_t chunk_z, int mode, double quality, size_t nthreads,
               void** dst, size_t* dst_len);
int so_decomp_3d(const void* src, size_t src_len, int output_float, size_t nthreads, size_t* dimx,
                 size_t* dimy, size_t* dimz, void** dst);
int so_comp_2d(const void* src, int is_float, size_t dimx, size_t dimy, int mode, double quality,
               int out_inc_header, void** dst, size_t* dst_len);
int so_decomp_2d(const void* src, size_t src_len, int output_float, size_t dimx, size_t dimy,
                 void** dst);

/* 0 (default): STRICT arithmetic; 1: the lifting steps contracted like the reference's stock x86 build */
void so_set_fma_flavour(int on);

int so_trunc_3d(const void* src, size_t src_len, unsigned pct, void** dst, size_t* dst_len);

#ifdef __cplusplus
}
#endif
#endif
