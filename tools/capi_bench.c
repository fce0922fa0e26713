/* A plain C caller of the reference-compatible API (include/SPERR_C_API.h): compresses and
 * decompresses a synthetic n^3 fp32 volume held in ordinary malloc'd memory, PWE mode, 256^3 chunks,
 * and prints the input GB/s of the round trip. The device set is chosen from outside
 * (SPERR_B200_DEVICES=all | 0,1,...; unset: one GPU), the program itself knows nothing about GPUs.
 *   cc -O2 tools/capi_bench.c -Iinclude -Lsperr_b200 -lsperr_b200 -Wl,-rpath,$PWD/sperr_b200 -lm -o tools/capi_bench
 *   tools/capi_bench [n=1024] [reps=3] [tol=1e-3]                                                   */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "SPERR_C_API.h"

static double now(void)
{
  struct timespec t;
  clock_gettime(CLOCK_MONOTONIC, &t);
  return (double)t.tv_sec + 1e-9 * (double)t.tv_nsec;
}

int main(int argc, char** argv)
{
  const size_t n = argc > 1 ? (size_t)atol(argv[1]) : 1024;
  const int reps = argc > 2 ? atoi(argv[2]) : 3;
  const double tol = argc > 3 ? atof(argv[3]) : 1e-3;
  const size_t total = n * n * n;
  float* vol = (float*)malloc(total * sizeof(float));
  if (!vol)
    return 1;
  /* smooth separable field: a few Fourier modes per axis */
  float* sx = (float*)malloc(n * 3 * sizeof(float));
  for (size_t i = 0; i < n; i++) {
    const double t = 6.283185307179586 * (double)i / 512.0;
    sx[i] = (float)(sin(3.1 * t + 0.3) + 0.5 * sin(7.7 * t + 1.1) + 0.25 * sin(19.3 * t + 2.0));
    sx[n + i] = (float)(sin(2.3 * t + 0.7) + 0.5 * sin(5.9 * t + 0.2) + 0.25 * sin(23.1 * t + 1.4));
    sx[2 * n + i] = (float)(sin(4.7 * t + 1.9) + 0.5 * sin(11.3 * t + 0.5) + 0.25 * sin(17.9 * t + 2.6));
  }
  for (size_t z = 0; z < n; z++)
    for (size_t y = 0; y < n; y++) {
      const float yz = sx[n + y] * sx[2 * n + z] * 0.2f;
      float* row = vol + (z * n + y) * n;
      for (size_t x = 0; x < n; x++)
        row[x] = sx[x] * yz;
    }
  double best = 1e30, best_c = 0, best_d = 0;
  size_t slen = 0;
  double maxerr = 0;
  for (int r = 0; r < reps + 1; r++) {   /* the first round trip is the warm-up */
    void* stream = NULL;
    size_t len = 0;
    const double t0 = now();
    int rc = sperr_comp_3d(vol, 1, n, n, n, 256, 256, 256, 3, tol, 0, &stream, &len);
    const double t1 = now();
    if (rc != 0) {
      fprintf(stderr, "sperr_comp_3d failed: %d\n", rc);
      return 2;
    }
    void* out = NULL;
    size_t dx = 0, dy = 0, dz = 0;
    rc = sperr_decomp_3d(stream, len, 1, 0, &dx, &dy, &dz, &out);
    const double t2 = now();
    if (rc != 0 || dx != n || dy != n || dz != n) {
      fprintf(stderr, "sperr_decomp_3d failed: %d\n", rc);
      return 3;
    }
    if (r == reps) {
      const float* o = (const float*)out;
      for (size_t i = 0; i < total; i += 7) {
        const double e = fabs((double)o[i] - (double)vol[i]);
        if (e > maxerr)
          maxerr = e;
      }
    }
    free(out);
    free(stream);
    slen = len;
    if (r > 0 && t2 - t0 < best) {
      best = t2 - t0;
      best_c = t1 - t0;
      best_d = t2 - t1;
    }
  }
  const char* devs = getenv("SPERR_B200_DEVICES");
  printf("{\"program\": \"capi_bench (C, malloc'd buffers)\", \"n\": %zu, \"devices\": \"%s\", \"round_trip_ms\": %.1f, "
         "\"compress_ms\": %.1f, \"decompress_ms\": %.1f, \"input_gbs\": %.2f, \"stream_bytes\": %zu, "
         "\"max_abs_err_sampled\": %.3g}\n",
         n, devs ? devs : "(one GPU)", best * 1e3, best_c * 1e3, best_d * 1e3, (double)total * 4.0 / best / 1e9, slen,
         maxerr);
  free(sx);
  free(vol);
  return 0;
}
