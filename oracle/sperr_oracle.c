/*
 * TEST INFRASTRUCTURE ONLY -- see sperr_oracle.h. Plain-C restatement of the SPERR v0.8.5 hot
 * path in STRICT arithmetic (compile with -ffp-contract=off). Citations are file:line under
 * /root/reference. Written for clarity, not speed.
 */
#define _GNU_SOURCE
#include "sperr_oracle.h"

#include <fenv.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* geometry                                                                                   */
/* ------------------------------------------------------------------------------------------ */

/* src/sperr_helper.cpp:36-49 */
size_t so_num_of_xforms(size_t len)
{
  size_t num = 0;
  while (len >= 9) {
    ++num;
    len -= len / 2;
  }
  return num < 6 ? num : 6;
}

/* src/sperr_helper.cpp:125-134 */
size_t so_num_of_partitions(size_t len)
{
  size_t n = 0;
  while (len > 1) {
    n++;
    len -= len / 2;
  }
  return n;
}

/* src/sperr_helper.cpp:51-68 */
int so_can_use_dyadic(size_t nx, size_t ny, size_t nz)
{
  if (nz < 2 || ny < 2)
    return -1;
  size_t xy = so_num_of_xforms(nx < ny ? nx : ny);
  size_t z = so_num_of_xforms(nz);
  if (xy == z || (xy >= 5 && z >= 5))
    return (int)(xy < z ? xy : z);
  return -1;
}

/* src/sperr_helper.cpp:136-146 */
void so_calc_approx_detail_len(size_t len, size_t lev, size_t out2[2])
{
  size_t low = len, high = 0;
  for (size_t i = 0; i < lev; i++) {
    high = low / 2;
    low -= high;
  }
  out2[0] = low;
  out2[1] = high;
}

/* src/sperr_helper.cpp:542-592 */
size_t so_chunk_volume(size_t vx, size_t vy, size_t vz, size_t cx, size_t cy, size_t cz,
                       size_t* out6, size_t cap)
{
  const size_t vol[3] = {vx, vy, vz}, chk[3] = {cx, cy, cz};
  size_t nseg[3];
  for (int i = 0; i < 3; i++) {
    nseg[i] = vol[i] / chk[i];
    if (vol[i] % chk[i] > chk[i] / 2)
      nseg[i]++;
    if (nseg[i] == 0)
      nseg[i] = 1;
  }
  size_t k = 0;
  for (size_t z = 0; z < nseg[2]; z++)
    for (size_t y = 0; y < nseg[1]; y++)
      for (size_t x = 0; x < nseg[0]; x++, k++) {
        if (k >= cap)
          continue;
        const size_t idx[3] = {x, y, z};
        for (int a = 0; a < 3; a++) {
          size_t beg = idx[a] * chk[a];
          size_t end = (idx[a] + 1 == nseg[a]) ? vol[a] : (idx[a] + 1) * chk[a];
          out6[k * 6 + 2 * a] = beg;
          out6[k * 6 + 2 * a + 1] = end - beg;
        }
      }
  return nseg[0] * nseg[1] * nseg[2];
}

/* ------------------------------------------------------------------------------------------ */
/* conditioner                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* src/Conditioner.cpp:137-163 */
static size_t adjust_strides(size_t len)
{
  size_t ns = 2048;
  if (len % ns == 0)
    return ns;
  size_t num;
  for (num = ns; num <= 32768; num++)
    if (len % num == 0)
      break;
  if (len % num == 0)
    return num;
  for (num = ns; num > 0; num--)
    if (len % num == 0)
      break;
  return num;
}

/* src/Conditioner.cpp:10-64, m_calc_mean :119-135. Header byte 0: pack_8_booleans puts
 * bool[0] in bit 7 (src/sperr_helper.cpp:262-273): 0x80 = mean subtracted, 0x01 = constant. */
int so_condition(double* buf, size_t n, uint8_t h[17])
{
  int constant = 1;
  for (size_t i = 0; i < n; i++)
    if (!(buf[i] == buf[0])) {
      constant = 0;
      break;
    }
  memset(h, 0, 17);
  if (constant) {
    h[0] = 0x81;
    uint64_t nval = n;
    memcpy(h + 1, &nval, 8);
    memcpy(h + 9, &buf[0], 8);
    return 1;
  }
  const size_t ns = adjust_strides(n);
  const size_t ss = n / ns;
  double sum = 0.0;
  for (size_t s = 0; s < ns; s++) {
    double acc = 0.0;
    for (size_t i = 0; i < ss; i++)
      acc += buf[s * ss + i];
    sum += acc / (double)ss;
  }
  const double mean = sum / (double)ns;
  for (size_t i = 0; i < n; i++)
    buf[i] -= mean;
  h[0] = 0x80;
  memcpy(h + 1, &mean, 8);
  return 0;
}

/* src/Conditioner.cpp:66-96 (constant fill is done by the caller, which knows n) */
void so_inverse_condition(double* buf, size_t n, const uint8_t h[17])
{
  if (h[0] & 0x01) {
    double val;
    memcpy(&val, h + 9, 8);
    for (size_t i = 0; i < n; i++)
      buf[i] = val;
    return;
  }
  double mean;
  memcpy(&mean, h + 1, 8);
  for (size_t i = 0; i < n; i++)
    buf[i] += mean;
}

/* ------------------------------------------------------------------------------------------ */
/* CDF 9/7                                                                                    */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  double ALPHA, BETA, GAMMA, DELTA, EPSILON, INV_EPSILON;
} cdf_consts;

/* include/CDF97.h:136-147 -- evaluated with the same expressions, not pasted. */
static cdf_consts cdf_get(void)
{
  const double h[5] = {0.602949018236, 0.266864118443, -0.078223266529, -0.016864118443,
                       0.026748757411};
  const double r0 = h[0] - 2.0 * h[4] * h[1] / h[3];
  const double r1 = h[2] - h[4] - h[4] * h[1] / h[3];
  const double s0 = h[1] - h[3] - h[3] * r0 / r1;
  const double t0 = h[0] - 2.0 * (h[2] - h[4]);
  cdf_consts c;
  c.ALPHA = h[4] / h[3];
  c.BETA = h[3] / r1;
  c.GAMMA = r1 / s0;
  c.DELTA = s0 / t0;
  c.EPSILON = sqrt(2.0) * t0;
  c.INV_EPSILON = 1.0 / c.EPSILON;
  return c;
}

/* Arithmetic flavour of the lifting steps. 0 = STRICT: every multiply and add rounded separately
 * (the reference built with -ffp-contract=off; what the CUDA kernels reproduce with -fmad=false).
 * 1 = FMA: the contraction pattern g++ 13 -O3 -mfma -ffp-contract=fast gives the reference's stock
 * x86 build (read from its disassembly): forward steps 1-3 fma(C, sum, x),
 * step 4 EPSILON * fma(DELTA, sum, e); inverse step 2 fma(e, INV_EPSILON, -(DELTA * sum)) with the
 * product rounded first, inverse steps 3-5 fma(-C, sum, x). */
static int g_fma_flavour = 0;
void so_set_fma_flavour(int on) { g_fma_flavour = on != 0; }

static void analysis_fma(const cdf_consts* c, double* sig, size_t len)
{
  const size_t el = len - len / 2, ol = len / 2;
  double *even = sig, *odd = sig + el;
  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] = fma(c->ALPHA, even[i] + even[i + 1], odd[i]);
  odd[ol - 1] = fma(c->ALPHA, even[ol - 1] + even[el - 1], odd[ol - 1]);

  even[0] = fma(2.0 * c->BETA, odd[0], even[0]);
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = fma(c->BETA, odd[i - 1] + odd[i], even[i]);
  even[el - 1] = fma(c->BETA, odd[el - 2] + odd[ol - 1], even[el - 1]);

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] = fma(c->GAMMA, even[i] + even[i + 1], odd[i]);
  odd[ol - 1] = fma(c->GAMMA, even[ol - 1] + even[el - 1], odd[ol - 1]);

  even[0] = c->EPSILON * fma(2.0 * c->DELTA, odd[0], even[0]);
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = c->EPSILON * fma(c->DELTA, odd[i - 1] + odd[i], even[i]);
  even[el - 1] = c->EPSILON * fma(c->DELTA, odd[el - 2] + odd[ol - 1], even[el - 1]);

  for (size_t i = 0; i < ol; i++)
    odd[i] *= -c->INV_EPSILON;
}

static void synthesis_fma(const cdf_consts* c, double* sig, size_t len)
{
  const size_t el = len - len / 2, ol = len / 2;
  double *even = sig, *odd = sig + el;
  for (size_t i = 0; i < ol; i++)
    odd[i] *= (-c->EPSILON);

  even[0] = fma(even[0], c->INV_EPSILON, -((2.0 * c->DELTA) * odd[0]));
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = fma(even[i], c->INV_EPSILON, -(c->DELTA * (odd[i - 1] + odd[i])));
  even[el - 1] = fma(even[el - 1], c->INV_EPSILON, -(c->DELTA * (odd[el - 2] + odd[ol - 1])));

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] = fma(-c->GAMMA, even[i] + even[i + 1], odd[i]);
  odd[ol - 1] = fma(-c->GAMMA, even[ol - 1] + even[el - 1], odd[ol - 1]);

  even[0] = fma(-(2.0 * c->BETA), odd[0], even[0]);
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = fma(-c->BETA, odd[i - 1] + odd[i], even[i]);
  even[el - 1] = fma(-c->BETA, odd[el - 2] + odd[ol - 1], even[el - 1]);

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] = fma(-c->ALPHA, even[i] + even[i + 1], odd[i]);
  odd[ol - 1] = fma(-c->ALPHA, even[ol - 1] + even[el - 1], odd[ol - 1]);
}

/* src/CDF97.cpp:598-631; `sig` holds evens then odds. */
static void analysis(const cdf_consts* c, double* sig, size_t len)
{
  if (g_fma_flavour) {
    analysis_fma(c, sig, len);
    return;
  }
  const size_t el = len - len / 2, ol = len / 2;
  double *even = sig, *odd = sig + el;
  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] += c->ALPHA * (even[i] + even[i + 1]);
  odd[ol - 1] += c->ALPHA * (even[ol - 1] + even[el - 1]);

  even[0] += 2.0 * c->BETA * odd[0];
  for (size_t i = 1; i + 1 < el; i++)
    even[i] += c->BETA * (odd[i - 1] + odd[i]);
  even[el - 1] += c->BETA * (odd[el - 2] + odd[ol - 1]);

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] += c->GAMMA * (even[i] + even[i + 1]);
  odd[ol - 1] += c->GAMMA * (even[ol - 1] + even[el - 1]);

  even[0] = c->EPSILON * (even[0] + 2.0 * c->DELTA * odd[0]);
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = c->EPSILON * (even[i] + c->DELTA * (odd[i - 1] + odd[i]));
  even[el - 1] = c->EPSILON * (even[el - 1] + c->DELTA * (odd[el - 2] + odd[ol - 1]));

  for (size_t i = 0; i < ol; i++)
    odd[i] *= -c->INV_EPSILON;
}

/* src/CDF97.cpp:633-666 */
static void synthesis(const cdf_consts* c, double* sig, size_t len)
{
  if (g_fma_flavour) {
    synthesis_fma(c, sig, len);
    return;
  }
  const size_t el = len - len / 2, ol = len / 2;
  double *even = sig, *odd = sig + el;
  for (size_t i = 0; i < ol; i++)
    odd[i] *= (-c->EPSILON);

  even[0] = even[0] * c->INV_EPSILON - 2.0 * c->DELTA * odd[0];
  for (size_t i = 1; i + 1 < el; i++)
    even[i] = even[i] * c->INV_EPSILON - c->DELTA * (odd[i - 1] + odd[i]);
  even[el - 1] = even[el - 1] * c->INV_EPSILON - c->DELTA * (odd[el - 2] + odd[ol - 1]);

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] -= c->GAMMA * (even[i] + even[i + 1]);
  odd[ol - 1] -= c->GAMMA * (even[ol - 1] + even[el - 1]);

  even[0] -= 2.0 * c->BETA * odd[0];
  for (size_t i = 1; i + 1 < el; i++)
    even[i] -= c->BETA * (odd[i - 1] + odd[i]);
  even[el - 1] -= c->BETA * (odd[el - 2] + odd[ol - 1]);

  for (size_t i = 0; i + 1 < ol; i++)
    odd[i] -= c->ALPHA * (even[i] + even[i + 1]);
  odd[ol - 1] -= c->ALPHA * (even[ol - 1] + even[el - 1]);
}

/* One forward 1D step on a strided line: m_gather (src/CDF97.cpp:476-519) + analysis. */
static void fwd_line(const cdf_consts* c, double* p, size_t stride, size_t len, double* tmp)
{
  const size_t el = len - len / 2;
  for (size_t i = 0; i < len; i++)
    tmp[(i & 1) ? el + i / 2 : i / 2] = p[i * stride];
  analysis(c, tmp, len);
  for (size_t i = 0; i < len; i++)
    p[i * stride] = tmp[i];
}

/* One inverse 1D step: synthesis + m_scatter (src/CDF97.cpp:521-564). */
static void inv_line(const cdf_consts* c, double* p, size_t stride, size_t len, double* tmp)
{
  const size_t el = len - len / 2;
  for (size_t i = 0; i < len; i++)
    tmp[i] = p[i * stride];
  synthesis(c, tmp, len);
  for (size_t i = 0; i < len; i++)
    p[i * stride] = tmp[(i & 1) ? el + i / 2 : i / 2];
}

/* src/CDF97.cpp:345-385; plane has row stride `sx`. */
static void dwt2d_one_level(const cdf_consts* c, double* plane, size_t sx, size_t lx, size_t ly,
                            double* tmp)
{
  for (size_t y = 0; y < ly; y++)
    fwd_line(c, plane + y * sx, 1, lx, tmp);
  for (size_t x = 0; x < lx; x++)
    fwd_line(c, plane + x, sx, ly, tmp);
}
static void idwt2d_one_level(const cdf_consts* c, double* plane, size_t sx, size_t lx, size_t ly,
                             double* tmp)
{
  for (size_t x = 0; x < lx; x++)
    inv_line(c, plane + x, sx, ly, tmp);
  for (size_t y = 0; y < ly; y++)
    inv_line(c, plane + y * sx, 1, lx, tmp);
}

/* src/CDF97.cpp:327-343 */
static void dwt2d_multi(const cdf_consts* c, double* plane, size_t nx, size_t ny, size_t nlev,
                        double* tmp)
{
  for (size_t lev = 0; lev < nlev; lev++) {
    size_t ax[2], ay[2];
    so_calc_approx_detail_len(nx, lev, ax);
    so_calc_approx_detail_len(ny, lev, ay);
    dwt2d_one_level(c, plane, nx, ax[0], ay[0], tmp);
  }
}
static void idwt2d_multi(const cdf_consts* c, double* plane, size_t nx, size_t ny, size_t nlev,
                         double* tmp)
{
  for (size_t lev = nlev; lev > 0; lev--) {
    size_t ax[2], ay[2];
    so_calc_approx_detail_len(nx, lev - 1, ax);
    so_calc_approx_detail_len(ny, lev - 1, ay);
    idwt2d_one_level(c, plane, nx, ax[0], ay[0], tmp);
  }
}

static size_t max3(size_t a, size_t b, size_t c)
{
  size_t m = a > b ? a : b;
  return m > c ? m : c;
}

/* src/CDF97.cpp:132-139, m_dwt3d_dyadic :284-292, m_dwt3d_one_level :387-429,
 * m_dwt3d_wavelet_packet :170-225. */
void so_dwt3d(double* buf, size_t nx, size_t ny, size_t nz)
{
  const cdf_consts c = cdf_get();
  double* tmp = (double*)malloc(sizeof(double) * max3(nx, ny, nz));
  const size_t plane = nx * ny;
  const int dy = so_can_use_dyadic(nx, ny, nz);
  if (dy >= 0) {
    for (size_t lev = 0; lev < (size_t)dy; lev++) {
      size_t ax[2], ay[2], az[2];
      so_calc_approx_detail_len(nx, lev, ax);
      so_calc_approx_detail_len(ny, lev, ay);
      so_calc_approx_detail_len(nz, lev, az);
      for (size_t z = 0; z < az[0]; z++)
        dwt2d_one_level(&c, buf + z * plane, nx, ax[0], ay[0], tmp);
      for (size_t y = 0; y < ay[0]; y++)
        for (size_t x = 0; x < ax[0]; x++)
          fwd_line(&c, buf + y * nx + x, plane, az[0], tmp);
    }
  }
  else {
    const size_t nz_x = so_num_of_xforms(nz);
    for (size_t y = 0; y < ny; y++)
      for (size_t x = 0; x < nx; x++) {
        size_t len = nz;
        for (size_t lev = 0; lev < nz_x; lev++) { /* m_dwt1d, src/CDF97.cpp:307-315 */
          fwd_line(&c, buf + y * nx + x, plane, len, tmp);
          len -= len / 2;
        }
      }
    const size_t nxy = so_num_of_xforms(nx < ny ? nx : ny);
    for (size_t z = 0; z < nz; z++)
      dwt2d_multi(&c, buf + z * plane, nx, ny, nxy, tmp);
  }
  free(tmp);
}

/* src/CDF97.cpp:141-148, m_idwt3d_dyadic :294-302, m_idwt3d_one_level :431-474,
 * m_idwt3d_wavelet_packet :227-282. */
void so_idwt3d(double* buf, size_t nx, size_t ny, size_t nz)
{
  const cdf_consts c = cdf_get();
  double* tmp = (double*)malloc(sizeof(double) * max3(nx, ny, nz));
  const size_t plane = nx * ny;
  const int dy = so_can_use_dyadic(nx, ny, nz);
  if (dy >= 0) {
    for (size_t lev = (size_t)dy; lev > 0; lev--) {
      size_t ax[2], ay[2], az[2];
      so_calc_approx_detail_len(nx, lev - 1, ax);
      so_calc_approx_detail_len(ny, lev - 1, ay);
      so_calc_approx_detail_len(nz, lev - 1, az);
      for (size_t y = 0; y < ay[0]; y++)
        for (size_t x = 0; x < ax[0]; x++)
          inv_line(&c, buf + y * nx + x, plane, az[0], tmp);
      for (size_t z = 0; z < az[0]; z++)
        idwt2d_one_level(&c, buf + z * plane, nx, ax[0], ay[0], tmp);
    }
  }
  else {
    const size_t nxy = so_num_of_xforms(nx < ny ? nx : ny);
    for (size_t z = 0; z < nz; z++)
      idwt2d_multi(&c, buf + z * plane, nx, ny, nxy, tmp);
    const size_t nz_x = so_num_of_xforms(nz);
    for (size_t y = 0; y < ny; y++)
      for (size_t x = 0; x < nx; x++)
        for (size_t lev = nz_x; lev > 0; lev--) { /* m_idwt1d, src/CDF97.cpp:317-325 */
          size_t a[2];
          so_calc_approx_detail_len(nz, lev - 1, a);
          inv_line(&c, buf + y * nx + x, plane, a[0], tmp);
        }
  }
  free(tmp);
}

/* src/CDF97.cpp:102-112 */
void so_dwt2d(double* buf, size_t nx, size_t ny)
{
  const cdf_consts c = cdf_get();
  double* tmp = (double*)malloc(sizeof(double) * (nx > ny ? nx : ny));
  dwt2d_multi(&c, buf, nx, ny, so_num_of_xforms(nx < ny ? nx : ny), tmp);
  free(tmp);
}
void so_idwt2d(double* buf, size_t nx, size_t ny)
{
  const cdf_consts c = cdf_get();
  double* tmp = (double*)malloc(sizeof(double) * (nx > ny ? nx : ny));
  idwt2d_multi(&c, buf, nx, ny, so_num_of_xforms(nx < ny ? nx : ny), tmp);
  free(tmp);
}

/* ------------------------------------------------------------------------------------------ */
/* quantiser                                                                                  */
/* ------------------------------------------------------------------------------------------ */

static int width_of(long long maxll)
{
  if (maxll <= 0xFF)
    return 1;
  if (maxll <= 0xFFFF)
    return 2;
  if (maxll <= 0xFFFFFFFFll)
    return 4;
  return 8;
}

/* src/SPECK_FLT.cpp:311-371 */
int so_quantize(const double* vals, size_t n, double q, uint64_t* mag, uint8_t* sign, int* width)
{
  fesetround(FE_TONEAREST);
  double maxd = vals[0]; /* std::max_element with |a| < |b|: first maximum */
  for (size_t i = 1; i < n; i++)
    if (fabs(maxd) < fabs(vals[i]))
      maxd = vals[i];
  feclearexcept(FE_INVALID);
  const long long maxll = llrint(fabs(maxd) / q);
  if (fetestexcept(FE_INVALID))
    return -1;
  *width = width_of(maxll);
  const double inv = 1.0 / q;
  for (size_t i = 0; i < n; i++) {
    const long long ll = llrint(vals[i] * inv);
    sign[i] = (ll >= 0);
    uint64_t m = (uint64_t)llabs(ll);
    /* the store truncates to the chosen integer type */
    if (*width == 1)
      m = (uint8_t)m;
    else if (*width == 2)
      m = (uint16_t)m;
    else if (*width == 4)
      m = (uint32_t)m;
    mag[i] = m;
  }
  return 0;
}

/* src/SPECK_FLT.cpp:373-399 */
void so_inv_quantize(const uint64_t* mag, const uint8_t* sign, size_t n, double q, double* vals)
{
  for (size_t i = 0; i < n; i++)
    vals[i] = q * (double)mag[i] * (sign[i] ? 1.0 : -1.0);
}

/* src/SPECK_FLT.cpp:237-266 (explicit fma, the only contraction in the STRICT flavour) */
double so_estimate_mse_midtread(const double* vals, size_t n, double q)
{
  const size_t stride = 4096, ns = n / stride;
  const double rcp = 1.0 / q;
  double total = 0.0;
  for (size_t s = 0; s <= ns; s++) {
    const size_t beg = s * stride, end = (s == ns) ? n : beg + stride;
    double acc = 0.0;
    for (size_t i = beg; i < end; i++) {
      const double d = fma(-q, rint(vals[i] * rcp), vals[i]);
      acc = acc + d * d;
    }
    total += acc;
  }
  return total / (double)n;
}

/* ------------------------------------------------------------------------------------------ */
/* bit I/O (src/Bitstream.cpp:74-173): bit k of the stream is bit (k & 7) of byte (k >> 3)     */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint8_t* buf;
  size_t cap; /* bytes */
  size_t n;   /* bits written */
} bitw;

static void bw_put(bitw* w, int bit)
{
  if ((w->n >> 3) >= w->cap) {
    size_t ncap = w->cap ? w->cap * 2 : 4096;
    w->buf = (uint8_t*)realloc(w->buf, ncap);
    memset(w->buf + w->cap, 0, ncap - w->cap);
    w->cap = ncap;
  }
  if (bit)
    w->buf[w->n >> 3] |= (uint8_t)(1u << (w->n & 7));
  w->n++;
}

typedef struct {
  const uint8_t* buf;
  size_t avail; /* bits really present; reads beyond return 0 (src/SPECK_INT.cpp:95-100) */
  size_t pos;
} bitr;

static int br_get(bitr* r)
{
  int b = 0;
  if (r->pos < r->avail)
    b = (r->buf[r->pos >> 3] >> (r->pos & 7)) & 1;
  r->pos++;
  return b;
}

/* ------------------------------------------------------------------------------------------ */
/* integer SPECK: shared state (src/SPECK_INT.cpp)                                            */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  uint64_t morton;
  uint32_t sx, sy, sz, lx, ly, lz; /* 1D uses sx/lx only; 2D uses x/y */
  uint32_t level;
} sset;

typedef struct {
  sset* a;
  size_t n, cap;
} slist;

static void sl_push(slist* l, sset s)
{
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 64;
    l->a = (sset*)realloc(l->a, l->cap * sizeof(sset));
  }
  l->a[l->n++] = s;
}

typedef struct {
  size_t* a;
  size_t n, cap;
} ivec;

static void iv_push(ivec* v, size_t x)
{
  if (v->n == v->cap) {
    v->cap = v->cap ? v->cap * 2 : 1024;
    v->a = (size_t*)realloc(v->a, v->cap * sizeof(size_t));
  }
  v->a[v->n++] = x;
}

typedef struct {
  int ndim; /* 1, 2 or 3 */
  int encoding;
  size_t nx, ny, nz, n;
  uint64_t* coeff; /* working magnitudes */
  uint8_t* sign;
  uint8_t* lip; /* one byte per coefficient */
  uint8_t* lsp;
  int8_t* msb; /* encoder: 3D = morton order, 2D = raster */
  ivec lsp_new;
  slist* lis;
  size_t nlis;
  sset I; /* 2D only */
  uint64_t thr;
  int msb_thr;
  bitw w;
  bitr r;
} speck;

static size_t s_nelem(const sset* s)
{
  return (size_t)s->lx * s->ly * s->lz;
}

static int msb_pos(uint64_t v)
{
  int p = -1;
  while (v) {
    v >>= 1;
    p++;
  }
  return p;
}

/* src/SPECK3D_INT.cpp:12-20 (and the 1D / 2D twins) */
static void clean_lis(speck* s)
{
  for (size_t l = 0; l < s->nlis; l++) {
    size_t k = 0;
    for (size_t i = 0; i < s->lis[l].n; i++)
      if (s->lis[l].a[i].lx != 0)
        s->lis[l].a[k++] = s->lis[l].a[i];
    s->lis[l].n = k;
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SPECK3D                                                                                    */
/* ------------------------------------------------------------------------------------------ */

/* src/SPECK3D_INT.cpp:214-326: returns the level of the children */
static uint32_t part_xyz(const sset* set, uint32_t lev, sset sub[8])
{
  const uint32_t spx[2] = {set->lx - set->lx / 2, set->lx / 2};
  const uint32_t spy[2] = {set->ly - set->ly / 2, set->ly / 2};
  const uint32_t spz[2] = {set->lz - set->lz / 2, set->lz / 2};
  lev += (spx[1] != 0) + (spy[1] != 0) + (spz[1] != 0);
  uint64_t mort = set->morton;
  for (int k = 0; k < 8; k++) {
    const int ix = k & 1, iy = (k >> 1) & 1, iz = (k >> 2) & 1;
    sset* c = &sub[k];
    c->morton = mort;
    c->sx = set->sx + (ix ? spx[0] : 0);
    c->sy = set->sy + (iy ? spy[0] : 0);
    c->sz = set->sz + (iz ? spz[0] : 0);
    c->lx = spx[ix];
    c->ly = spy[iy];
    c->lz = spz[iz];
    c->level = lev;
    mort += s_nelem(c);
  }
  return lev;
}

/* src/SPECK3D_INT.cpp:22-97 */
static void s3_init_lists(speck* s)
{
  s->nlis = so_num_of_partitions(s->nx) + so_num_of_partitions(s->ny) +
            so_num_of_partitions(s->nz) + 1;
  s->lis = (slist*)calloc(s->nlis, sizeof(slist));
  sset big;
  memset(&big, 0, sizeof(big));
  big.lx = (uint32_t)s->nx;
  big.ly = (uint32_t)s->ny;
  big.lz = (uint32_t)s->nz;
  uint32_t cur = 0;
  sset sub[8];
  const int dy = so_can_use_dyadic(s->nx, s->ny, s->nz);
  size_t nxy, nzx, xf = 0;
  if (dy >= 0) {
    nxy = nzx = (size_t)dy;
  }
  else {
    nxy = so_num_of_xforms(s->nx < s->ny ? s->nx : s->ny);
    nzx = so_num_of_xforms(s->nz);
  }
  while (xf < nxy && xf < nzx) {
    uint32_t nl = part_xyz(&big, cur, sub);
    big = sub[0];
    for (int k = 1; k < 8; k++)
      sl_push(&s->lis[nl], sub[k]);
    cur = nl;
    xf++;
  }
  while (xf < nxy) { /* m_partition_S_XY, src/SPECK3D_INT.cpp:328-392 */
    const uint32_t spx[2] = {big.lx - big.lx / 2, big.lx / 2};
    const uint32_t spy[2] = {big.ly - big.ly / 2, big.ly / 2};
    uint32_t nl = cur + (spx[1] != 0) + (spy[1] != 0);
    for (int k = 1; k < 4; k++) {
      sset c = big;
      const int ix = k & 1, iy = k >> 1;
      c.sx = big.sx + (ix ? spx[0] : 0);
      c.sy = big.sy + (iy ? spy[0] : 0);
      c.lx = spx[ix];
      c.ly = spy[iy];
      c.level = nl;
      sl_push(&s->lis[nl], c);
    }
    big.lx = spx[0];
    big.ly = spy[0];
    cur = nl;
    xf++;
  }
  while (xf < nzx) { /* m_partition_S_Z, src/SPECK3D_INT.cpp:394-430 */
    const uint32_t spz[2] = {big.lz - big.lz / 2, big.lz / 2};
    uint32_t nl = cur + (spz[1] != 0);
    sset c = big;
    c.sz = big.sz + spz[0];
    c.lz = spz[1];
    c.level = nl;
    sl_push(&s->lis[nl], c);
    big.lz = spz[0];
    cur = nl;
    xf++;
  }
  /* insert `big` at the front of its list (src/SPECK3D_INT.cpp:93) */
  big.level = cur;
  sl_push(&s->lis[cur], big);
  slist* l = &s->lis[cur];
  memmove(l->a + 1, l->a, (l->n - 1) * sizeof(sset));
  l->a[0] = big;
}

/* src/SPECK3D_INT_ENC.cpp:8-139 (the unrolled small cases are equivalent to the recursion) */
static void s3_deposit(speck* s, const sset* set)
{
  const size_t ne = s_nelem(set);
  if (ne == 0)
    return;
  if (ne == 1) {
    size_t id = (size_t)set->sz * s->nx * s->ny + (size_t)set->sy * s->nx + set->sx;
    s->msb[set->morton] = (int8_t)msb_pos(s->coeff[id]);
    return;
  }
  sset sub[8];
  part_xyz(set, 0, sub);
  for (int k = 0; k < 8; k++)
    s3_deposit(s, &sub[k]);
}

/* src/SPECK3D_INT_ENC.cpp:141-159 */
static void s3_enc_additional_init(speck* s)
{
  s->msb = (int8_t*)malloc(s->n);
  uint64_t off = 0;
  for (size_t t = 1; t <= s->nlis; t++) {
    slist* l = &s->lis[s->nlis - t];
    for (size_t i = 0; i < l->n; i++) {
      l->a[i].morton = off;
      s3_deposit(s, &l->a[i]);
      off += s_nelem(&l->a[i]);
    }
  }
}

static size_t s3_process_S(speck* s, size_t i1, size_t i2, int need);

/* m_process_P: src/SPECK3D_INT_ENC.cpp:182-198, src/SPECK3D_INT_DEC.cpp:24-37 */
static size_t s3_process_P(speck* s, size_t idx, uint64_t morton, int need)
{
  int sig = 1;
  if (s->encoding) {
    if (need) {
      sig = (s->msb[morton] >= s->msb_thr);
      bw_put(&s->w, sig);
    }
    if (sig)
      bw_put(&s->w, s->sign[idx]);
  }
  else {
    if (need)
      sig = br_get(&s->r);
    if (sig)
      s->sign[idx] = (uint8_t)br_get(&s->r);
  }
  if (sig) {
    iv_push(&s->lsp_new, idx);
    s->lip[idx] = 0;
  }
  return (size_t)sig;
}

/* src/SPECK3D_INT.cpp:140-212 */
static void s3_code_S(speck* s, size_t i1, size_t i2)
{
  const sset set = s->lis[i1].a[i2];
  sset sub[8];
  const uint32_t nl = part_xyz(&set, (uint32_t)i1, sub);
  sset* ne[8];
  int cnt = 0;
  for (int k = 0; k < 8; k++)
    if (s_nelem(&sub[k]) != 0)
      ne[cnt++] = &sub[k];
  size_t sigc = 0;
  for (int k = 0; k < cnt; k++) {
    const int need = (sigc != 0 || k + 1 != cnt);
    sset* c = ne[k];
    if (s_nelem(c) == 1) {
      size_t idx = (size_t)c->sz * s->nx * s->ny + (size_t)c->sy * s->nx + c->sx;
      s->lip[idx] = 1;
      sigc += s3_process_P(s, idx, c->morton, need);
    }
    else {
      sl_push(&s->lis[nl], *c);
      sigc += s3_process_S(s, nl, s->lis[nl].n - 1, need);
    }
  }
}

/* src/SPECK3D_INT_ENC.cpp:161-180, src/SPECK3D_INT_DEC.cpp:8-22 */
static size_t s3_process_S(speck* s, size_t i1, size_t i2, int need)
{
  int sig = 1;
  if (need) {
    if (s->encoding) {
      const sset* set = &s->lis[i1].a[i2];
      const size_t ne = s_nelem(set);
      sig = 0;
      for (size_t k = 0; k < ne; k++)
        if (s->msb[set->morton + k] >= s->msb_thr) {
          sig = 1;
          break;
        }
      bw_put(&s->w, sig);
    }
    else
      sig = br_get(&s->r);
  }
  if (sig) {
    s3_code_S(s, i1, i2);
    s->lis[i1].a[i2].lx = 0; /* make_empty */
  }
  return (size_t)sig;
}

/* src/SPECK3D_INT.cpp:99-138; m_process_P_lite src/SPECK3D_INT_ENC.cpp:200-212, _DEC:39-49 */
static void s3_sorting_pass(speck* s)
{
  for (size_t i = 0; i < s->n; i++) {
    if (!s->lip[i])
      continue;
    int sig;
    if (s->encoding) {
      sig = (s->coeff[i] >= s->thr);
      bw_put(&s->w, sig);
      if (sig)
        bw_put(&s->w, s->sign[i]);
    }
    else {
      sig = br_get(&s->r);
      if (sig)
        s->sign[i] = (uint8_t)br_get(&s->r);
    }
    if (sig) {
      iv_push(&s->lsp_new, i);
      s->lip[i] = 0;
    }
  }
  for (size_t t = 1; t <= s->nlis; t++) {
    const size_t i1 = s->nlis - t;
    for (size_t i2 = 0; i2 < s->lis[i1].n; i2++)
      s3_process_S(s, i1, i2, 1);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SPECK1D (src/SPECK1D_INT.cpp, _ENC.cpp, _DEC.cpp)                                          */
/* ------------------------------------------------------------------------------------------ */

/* For 1D the fields of `sset` are too narrow for 64-bit lengths, so use a dedicated type. */
typedef struct {
  uint64_t start, len;
  uint32_t level;
} set1;
typedef struct {
  set1* a;
  size_t n, cap;
} list1;
static void l1_push(list1* l, set1 s)
{
  if (l->n == l->cap) {
    l->cap = l->cap ? l->cap * 2 : 64;
    l->a = (set1*)realloc(l->a, l->cap * sizeof(set1));
  }
  l->a[l->n++] = s;
}

typedef struct {
  int encoding;
  size_t n;
  uint64_t* coeff;
  uint8_t *sign, *lip, *lsp;
  ivec lsp_new;
  list1* lis;
  size_t nlis;
  uint64_t thr;
  bitw w;
  bitr r;
} speck1;

static size_t s1_process_S(speck1* s, size_t i1, size_t i2, int need);

/* src/SPECK1D_INT_ENC.cpp:95-119 (encoder subtracts the threshold at discovery),
 * src/SPECK1D_INT_DEC.cpp:73-90 */
static size_t s1_process_P(speck1* s, size_t idx, int need)
{
  int sig = 1;
  if (s->encoding) {
    /* SigType bookkeeping in the reference only avoids re-testing; the emitted bit is the
     * plain significance of the pixel. When !need the pixel is known significant. */
    if (need) {
      sig = (s->coeff[idx] >= s->thr);
      bw_put(&s->w, sig);
    }
    if (sig) {
      bw_put(&s->w, s->sign[idx]);
      s->coeff[idx] -= s->thr;
    }
  }
  else {
    if (need)
      sig = br_get(&s->r);
    if (sig)
      s->sign[idx] = (uint8_t)br_get(&s->r);
  }
  if (sig) {
    iv_push(&s->lsp_new, idx);
    s->lip[idx] = 0;
  }
  return (size_t)sig;
}

/* src/SPECK1D_INT_ENC.cpp:121-160, src/SPECK1D_INT_DEC.cpp:92-125; partition
 * src/SPECK1D_INT.cpp:36-56 */
static void s1_code_S(speck1* s, size_t i1, size_t i2)
{
  const set1 set = s->lis[i1].a[i2];
  set1 sub[2];
  sub[0].start = set.start;
  sub[0].len = set.len - set.len / 2;
  sub[0].level = set.level + 1;
  sub[1].start = set.start + set.len - set.len / 2;
  sub[1].len = set.len / 2;
  sub[1].level = set.level + 1;
  size_t sigc = 0;
  for (int k = 0; k < 2; k++) {
    const int need = (k == 0) ? 1 : (sigc != 0);
    if (sub[k].len == 1) {
      s->lip[sub[k].start] = 1;
      sigc += s1_process_P(s, sub[k].start, need);
    }
    else {
      l1_push(&s->lis[sub[k].level], sub[k]);
      sigc += s1_process_S(s, sub[k].level, s->lis[sub[k].level].n - 1, need);
    }
  }
}

/* src/SPECK1D_INT_ENC.cpp:55-93, src/SPECK1D_INT_DEC.cpp:56-71 */
static size_t s1_process_S(speck1* s, size_t i1, size_t i2, int need)
{
  int sig = 1;
  if (s->encoding) {
    if (need) {
      const set1* set = &s->lis[i1].a[i2];
      sig = 0;
      for (uint64_t k = 0; k < set->len; k++)
        if (s->coeff[set->start + k] >= s->thr) {
          sig = 1;
          break;
        }
      bw_put(&s->w, sig);
    }
  }
  else if (need)
    sig = br_get(&s->r);
  if (sig) {
    s1_code_S(s, i1, i2);
    s->lis[i1].a[i2].len = 0;
  }
  return (size_t)sig;
}

static void s1_sorting_pass(speck1* s)
{
  for (size_t i = 0; i < s->n; i++)
    if (s->lip[i])
      s1_process_P(s, i, 1);
  for (size_t t = 1; t <= s->nlis; t++) {
    const size_t i1 = s->nlis - t;
    for (size_t i2 = 0; i2 < s->lis[i1].n; i2++)
      s1_process_S(s, i1, i2, 1);
  }
}

/* ------------------------------------------------------------------------------------------ */
/* SPECK2D (src/SPECK2D_INT.cpp, _ENC.cpp, _DEC.cpp)                                          */
/* ------------------------------------------------------------------------------------------ */

static size_t s2_process_S(speck* s, size_t i1, size_t i2, int need);
static void s2_process_I(speck* s, int need);

/* src/SPECK2D_INT_ENC.cpp:27-42, src/SPECK2D_INT_DEC.cpp */
static size_t s2_process_P(speck* s, size_t idx, int need)
{
  int sig = 1;
  if (s->encoding) {
    if (need) {
      sig = (s->msb[idx] >= s->msb_thr);
      bw_put(&s->w, sig);
    }
    if (sig)
      bw_put(&s->w, s->sign[idx]);
  }
  else {
    if (need)
      sig = br_get(&s->r);
    if (sig)
      s->sign[idx] = (uint8_t)br_get(&s->r);
  }
  if (sig) {
    iv_push(&s->lsp_new, idx);
    s->lip[idx] = 0;
  }
  return (size_t)sig;
}

/* src/SPECK2D_INT.cpp:109-148: children in the order BR, BL, TR, TL */
static void s2_partition_S(const sset* set, sset sub[4])
{
  const uint32_t dx = set->lx / 2, dy = set->ly / 2;
  const uint32_t ax = set->lx - dx, ay = set->ly - dy;
  for (int k = 0; k < 4; k++) {
    sub[k] = *set;
    sub[k].level = set->level + 1;
  }
  sub[0].sx = set->sx + ax; sub[0].sy = set->sy + ay; sub[0].lx = dx; sub[0].ly = dy;
  sub[1].sx = set->sx;      sub[1].sy = set->sy + ay; sub[1].lx = ax; sub[1].ly = dy;
  sub[2].sx = set->sx + ax; sub[2].sy = set->sy;      sub[2].lx = dx; sub[2].ly = ay;
  sub[3].sx = set->sx;      sub[3].sy = set->sy;      sub[3].lx = ax; sub[3].ly = ay;
}

/* src/SPECK2D_INT.cpp:57-80 */
static void s2_code_S(speck* s, size_t i1, size_t i2)
{
  const sset set = s->lis[i1].a[i2];
  sset sub[4], *ne[4];
  s2_partition_S(&set, sub);
  int cnt = 0;
  for (int k = 0; k < 4; k++)
    if ((size_t)sub[k].lx * sub[k].ly != 0)
      ne[cnt++] = &sub[k];
  size_t sigc = 0;
  for (int k = 0; k < cnt; k++) {
    const int need = (sigc != 0) || (k + 1 != cnt);
    sset* c = ne[k];
    if ((size_t)c->lx * c->ly == 1) {
      size_t idx = (size_t)c->sy * s->nx + c->sx;
      s->lip[idx] = 1;
      sigc += s2_process_P(s, idx, need);
    }
    else {
      sl_push(&s->lis[c->level], *c);
      sigc += s2_process_S(s, c->level, s->lis[c->level].n - 1, need);
    }
  }
}

/* src/SPECK2D_INT_ENC.cpp:7-25, m_decide_S_significance :57-68 */
static size_t s2_process_S(speck* s, size_t i1, size_t i2, int need)
{
  int sig = 1;
  if (need) {
    if (s->encoding) {
      const sset* set = &s->lis[i1].a[i2];
      sig = 0;
      for (uint32_t y = set->sy; y < set->sy + set->ly && !sig; y++)
        for (uint32_t x = set->sx; x < set->sx + set->lx; x++)
          if (s->msb[(size_t)y * s->nx + x] >= s->msb_thr) {
            sig = 1;
            break;
          }
      bw_put(&s->w, sig);
    }
    else
      sig = br_get(&s->r);
  }
  if (sig) {
    s2_code_S(s, i1, i2);
    s->lis[i1].a[i2].lx = 0;
  }
  return (size_t)sig;
}

/* src/SPECK2D_INT.cpp:82-98 with m_partition_I :150-185 */
static void s2_code_I(speck* s)
{
  size_t ax[2], ay[2];
  so_calc_approx_detail_len(s->nx, s->I.level, ax);
  so_calc_approx_detail_len(s->ny, s->I.level, ay);
  sset sub[3];
  memset(sub, 0, sizeof(sub));
  for (int k = 0; k < 3; k++) {
    sub[k].lz = 1;
    sub[k].level = s->I.level;
  }
  sub[0].sx = (uint32_t)ax[0]; sub[0].sy = (uint32_t)ay[0]; sub[0].lx = (uint32_t)ax[1]; sub[0].ly = (uint32_t)ay[1];
  sub[1].sx = (uint32_t)ax[0]; sub[1].sy = 0;               sub[1].lx = (uint32_t)ax[1]; sub[1].ly = (uint32_t)ay[0];
  sub[2].sx = 0;               sub[2].sy = (uint32_t)ay[0]; sub[2].lx = (uint32_t)ax[0]; sub[2].ly = (uint32_t)ay[1];
  s->I.sx += (uint32_t)ax[1];
  s->I.sy += (uint32_t)ay[1];
  s->I.level--;
  size_t sigc = 0;
  for (int k = 0; k < 3; k++)
    if ((size_t)sub[k].lx * sub[k].ly != 0) {
      sl_push(&s->lis[sub[k].level], sub[k]);
      sigc += s2_process_S(s, sub[k].level, s->lis[sub[k].level].n - 1, 1);
    }
  s2_process_I(s, sigc != 0);
}

/* src/SPECK2D_INT_ENC.cpp:44-55, m_decide_I_significance :70-94 */
static void s2_process_I(speck* s, int need)
{
  if (s->I.level == 0)
    return;
  int sig = 1;
  if (need) {
    if (s->encoding) {
      sig = 0;
      for (size_t y = 0; y < s->ny && !sig; y++)
        for (size_t x = 0; x < s->nx; x++) {
          if (y < s->I.sy && x < s->I.sx)
            continue;
          if (s->msb[y * s->nx + x] >= s->msb_thr) {
            sig = 1;
            break;
          }
        }
      bw_put(&s->w, sig);
    }
    else
      sig = br_get(&s->r);
  }
  if (sig)
    s2_code_I(s);
}

/* src/SPECK2D_INT.cpp:187-218 */
static void s2_init_lists(speck* s)
{
  const size_t mx = s->nx > s->ny ? s->nx : s->ny;
  s->nlis = so_num_of_partitions(mx) + 1;
  const size_t nxf = so_num_of_xforms(s->nx < s->ny ? s->nx : s->ny);
  if (s->nlis < nxf + 1)
    s->nlis = nxf + 1;
  s->lis = (slist*)calloc(s->nlis, sizeof(slist));
  size_t ax[2], ay[2];
  so_calc_approx_detail_len(s->nx, nxf, ax);
  so_calc_approx_detail_len(s->ny, nxf, ay);
  sset root;
  memset(&root, 0, sizeof(root));
  root.lx = (uint32_t)ax[0];
  root.ly = (uint32_t)ay[0];
  root.lz = 1;
  root.level = (uint32_t)nxf;
  sl_push(&s->lis[nxf], root);
  memset(&s->I, 0, sizeof(s->I));
  s->I.sx = root.lx;
  s->I.sy = root.ly;
  s->I.lx = (uint32_t)s->nx;
  s->I.ly = (uint32_t)s->ny;
  s->I.lz = 1;
  s->I.level = (uint32_t)nxf;
}

static void s2_sorting_pass(speck* s)
{
  for (size_t i = 0; i < s->n; i++)
    if (s->lip[i])
      s2_process_P(s, i, 1);
  for (size_t t = 1; t <= s->nlis; t++) {
    const size_t i1 = s->nlis - t;
    for (size_t i2 = 0; i2 < s->lis[i1].n; i2++)
      s2_process_S(s, i1, i2, 1);
  }
  s2_process_I(s, 1);
}

/* ------------------------------------------------------------------------------------------ */
/* SPECK_INT driver: encode (src/SPECK_INT.cpp:110-163), decode (:165-228), refinement         */
/* passes (:310-469), stream header (:79-108, :264-308)                                        */
/* ------------------------------------------------------------------------------------------ */

static size_t finish_stream(uint8_t planes, const bitw* w, size_t budget_bits, uint8_t* out,
                            size_t cap)
{
  const uint64_t total = w->n;
  size_t pack = total;
  if (budget_bits != 0 && budget_bits < pack)
    pack = budget_bits;
  const size_t nbytes = (pack + 7) / 8;
  if (9 + nbytes <= cap) {
    out[0] = planes;
    memcpy(out + 1, &total, 8);
    if (nbytes)
      memcpy(out + 9, w->buf, nbytes);
  }
  return 9 + nbytes;
}

static size_t round_budget(size_t b)
{
  if (b == 0)
    return 0;
  while (b % 8 != 0)
    b++;
  return b;
}

/* shared refinement passes over raster order */
static void refine_encode(uint64_t* coeff, uint8_t* lsp, size_t n, uint64_t thr, bitw* w,
                          ivec* lsp_new, int subtract_new)
{
  for (size_t i = 0; i < n; i++)
    if (lsp[i]) {
      const int o = coeff[i] >= thr;
      if (o)
        coeff[i] -= thr;
      bw_put(w, o);
    }
  for (size_t k = 0; k < lsp_new->n; k++) {
    if (subtract_new) /* m_refinement_extra: 3D and 2D only */
      coeff[lsp_new->a[k]] -= thr;
    lsp[lsp_new->a[k]] = 1;
  }
  lsp_new->n = 0;
}

static void refine_decode(uint64_t* coeff, uint8_t* lsp, size_t n, uint64_t thr, bitr* r,
                          ivec* lsp_new)
{
  const uint64_t half = thr / 2;
  for (size_t i = 0; i < n; i++)
    if (lsp[i]) {
      const int b = br_get(r);
      if (thr >= 2) {
        if (b)
          coeff[i] += half;
        else
          coeff[i] -= half;
      }
      else if (b)
        coeff[i]++;
      if (r->pos == r->avail)
        break;
    }
  const uint64_t init = thr + thr - thr / 2 - 1;
  for (size_t k = 0; k < lsp_new->n; k++) {
    coeff[lsp_new->a[k]] = init;
    lsp[lsp_new->a[k]] = 1;
  }
  lsp_new->n = 0;
}

static void speck_free(speck* s)
{
  for (size_t l = 0; l < s->nlis; l++)
    free(s->lis[l].a);
  free(s->lis);
  free(s->lip);
  free(s->lsp);
  free(s->msb);
  free(s->lsp_new.a);
}

static size_t speck_nd_encode(int ndim, const uint64_t* mag, const uint8_t* sign, size_t nx,
                              size_t ny, size_t nz, size_t budget_bits, uint8_t* out, size_t cap)
{
  speck s;
  memset(&s, 0, sizeof(s));
  s.ndim = ndim;
  s.encoding = 1;
  s.nx = nx; s.ny = ny; s.nz = nz;
  s.n = nx * ny * nz;
  const size_t budget = round_budget(budget_bits);
  s.coeff = (uint64_t*)malloc(s.n * 8);
  memcpy(s.coeff, mag, s.n * 8);
  s.sign = (uint8_t*)sign;
  s.lip = (uint8_t*)calloc(s.n, 1);
  s.lsp = (uint8_t*)calloc(s.n, 1);
  if (ndim == 3) {
    s3_init_lists(&s);
    s3_enc_additional_init(&s);
  }
  else {
    s2_init_lists(&s);
    s.msb = (int8_t*)malloc(s.n);
    for (size_t i = 0; i < s.n; i++)
      s.msb[i] = (int8_t)msb_pos(s.coeff[i]);
  }
  uint64_t maxc = 0;
  for (size_t i = 0; i < s.n; i++)
    if (s.coeff[i] > maxc)
      maxc = s.coeff[i];
  uint8_t planes = 0;
  if (maxc != 0) {
    planes = 1;
    s.thr = 1;
    while (maxc - s.thr >= s.thr) {
      s.thr *= 2;
      planes++;
    }
    for (uint8_t bp = 0; bp < planes; bp++) {
      s.msb_thr = msb_pos(s.thr);
      if (ndim == 3)
        s3_sorting_pass(&s);
      else
        s2_sorting_pass(&s);
      if (budget && s.w.n >= budget)
        break;
      refine_encode(s.coeff, s.lsp, s.n, s.thr, &s.w, &s.lsp_new, 1);
      if (budget && s.w.n >= budget)
        break;
      s.thr /= 2;
      clean_lis(&s);
    }
  }
  size_t len = finish_stream(planes, &s.w, budget, out, cap);
  free(s.w.buf);
  free(s.coeff);
  speck_free(&s);
  return len;
}

static void speck_nd_decode(int ndim, const uint8_t* stream, size_t len, size_t nx, size_t ny,
                            size_t nz, uint64_t* mag, uint8_t* sign)
{
  speck s;
  memset(&s, 0, sizeof(s));
  s.ndim = ndim;
  s.nx = nx; s.ny = ny; s.nz = nz;
  s.n = nx * ny * nz;
  s.coeff = mag;
  s.sign = sign;
  memset(mag, 0, s.n * 8);
  memset(sign, 1, s.n);
  const uint8_t planes = stream[0];
  uint64_t total;
  memcpy(&total, stream + 1, 8);
  size_t avail = (len - 9) * 8;
  if (avail > total)
    avail = total;
  s.r.buf = stream + 9;
  s.r.avail = avail;
  s.r.pos = 0;
  if (planes == 0)
    return;
  s.lip = (uint8_t*)calloc(s.n, 1);
  s.lsp = (uint8_t*)calloc(s.n, 1);
  if (ndim == 3)
    s3_init_lists(&s);
  else
    s2_init_lists(&s);
  s.thr = 1;
  for (uint8_t i = 1; i < planes; i++)
    s.thr *= 2;
  for (uint8_t bp = 0; bp < planes; bp++) {
    if (ndim == 3)
      s3_sorting_pass(&s);
    else
      s2_sorting_pass(&s);
    if (s.r.pos >= avail)
      break;
    refine_decode(s.coeff, s.lsp, s.n, s.thr, &s.r, &s.lsp_new);
    if (s.r.pos >= avail)
      break;
    s.thr /= 2;
    clean_lis(&s);
  }
  if (s.lsp_new.n) {
    const uint64_t init = s.thr + s.thr - s.thr / 2 - 1;
    for (size_t k = 0; k < s.lsp_new.n; k++)
      s.coeff[s.lsp_new.a[k]] = init;
  }
  speck_free(&s);
}

size_t so_speck3d_encode(const uint64_t* mag, const uint8_t* sign, size_t nx, size_t ny, size_t nz,
                         size_t budget_bits, uint8_t* out, size_t cap)
{
  return speck_nd_encode(3, mag, sign, nx, ny, nz, budget_bits, out, cap);
}
void so_speck3d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, size_t nz,
                       uint64_t* mag, uint8_t* sign)
{
  speck_nd_decode(3, stream, len, nx, ny, nz, mag, sign);
}
size_t so_speck2d_encode(const uint64_t* mag, const uint8_t* sign, size_t nx, size_t ny,
                         size_t budget_bits, uint8_t* out, size_t cap)
{
  return speck_nd_encode(2, mag, sign, nx, ny, 1, budget_bits, out, cap);
}
void so_speck2d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, uint64_t* mag,
                       uint8_t* sign)
{
  speck_nd_decode(2, stream, len, nx, ny, 1, mag, sign);
}

/* src/SPECK1D_INT.cpp:18-34 */
static void s1_setup_lists(speck1* s)
{
  s->nlis = so_num_of_partitions(s->n) + 1;
  s->lis = (list1*)calloc(s->nlis + 1, sizeof(list1));
  set1 a, b;
  a.start = 0;
  a.len = s->n - s->n / 2;
  a.level = 1;
  b.start = s->n - s->n / 2;
  b.len = s->n / 2;
  b.level = 1;
  l1_push(&s->lis[1], a);
  l1_push(&s->lis[1], b);
}

static void s1_clean(speck1* s)
{
  for (size_t l = 0; l <= s->nlis; l++) {
    size_t k = 0;
    for (size_t i = 0; i < s->lis[l].n; i++)
      if (s->lis[l].a[i].len != 0)
        s->lis[l].a[k++] = s->lis[l].a[i];
    s->lis[l].n = k;
  }
}

static void s1_free(speck1* s)
{
  for (size_t l = 0; l <= s->nlis; l++)
    free(s->lis[l].a);
  free(s->lis);
  free(s->lip);
  free(s->lsp);
  free(s->lsp_new.a);
}

size_t so_speck1d_encode(const uint64_t* mag, const uint8_t* sign, size_t n, uint8_t* out,
                         size_t cap)
{
  speck1 s;
  memset(&s, 0, sizeof(s));
  s.encoding = 1;
  s.n = n;
  s.coeff = (uint64_t*)malloc(n * 8);
  memcpy(s.coeff, mag, n * 8);
  s.sign = (uint8_t*)sign;
  s.lip = (uint8_t*)calloc(n, 1);
  s.lsp = (uint8_t*)calloc(n, 1);
  s1_setup_lists(&s);
  uint64_t maxc = 0;
  for (size_t i = 0; i < n; i++)
    if (s.coeff[i] > maxc)
      maxc = s.coeff[i];
  uint8_t planes = 0;
  if (maxc != 0) {
    planes = 1;
    s.thr = 1;
    while (maxc - s.thr >= s.thr) {
      s.thr *= 2;
      planes++;
    }
    for (uint8_t bp = 0; bp < planes; bp++) {
      s1_sorting_pass(&s);
      refine_encode(s.coeff, s.lsp, n, s.thr, &s.w, &s.lsp_new, 0);
      s.thr /= 2;
      s1_clean(&s);
    }
  }
  size_t len = finish_stream(planes, &s.w, 0, out, cap);
  free(s.w.buf);
  free(s.coeff);
  s1_free(&s);
  return len;
}

void so_speck1d_decode(const uint8_t* stream, size_t len, size_t n, uint64_t* mag, uint8_t* sign)
{
  speck1 s;
  memset(&s, 0, sizeof(s));
  s.n = n;
  s.coeff = mag;
  s.sign = sign;
  memset(mag, 0, n * 8);
  memset(sign, 1, n);
  const uint8_t planes = stream[0];
  uint64_t total;
  memcpy(&total, stream + 1, 8);
  size_t avail = (len - 9) * 8;
  if (avail > total)
    avail = total;
  s.r.buf = stream + 9;
  s.r.avail = avail;
  if (planes == 0)
    return;
  s.lip = (uint8_t*)calloc(n, 1);
  s.lsp = (uint8_t*)calloc(n, 1);
  s1_setup_lists(&s);
  s.thr = 1;
  for (uint8_t i = 1; i < planes; i++)
    s.thr *= 2;
  for (uint8_t bp = 0; bp < planes; bp++) {
    s1_sorting_pass(&s);
    if (s.r.pos >= avail)
      break;
    refine_decode(s.coeff, s.lsp, n, s.thr, &s.r, &s.lsp_new);
    if (s.r.pos >= avail)
      break;
    s.thr /= 2;
    s1_clean(&s);
  }
  if (s.lsp_new.n) {
    const uint64_t init = s.thr + s.thr - s.thr / 2 - 1;
    for (size_t k = 0; k < s.lsp_new.n; k++)
      s.coeff[s.lsp_new.a[k]] = init;
  }
  s1_free(&s);
}

/* ------------------------------------------------------------------------------------------ */
/* outlier coder (src/Outlier_Coder.cpp)                                                      */
/* ------------------------------------------------------------------------------------------ */

/* encode :71-131, m_quantize :188-204 */
size_t so_outlier_encode(const uint64_t* pos, const double* err, size_t n_out, size_t total_len,
                         double tol, uint8_t* out, size_t cap)
{
  if (total_len == 0 || tol <= 0.0 || n_out == 0)
    return 0;
  for (size_t i = 0; i < n_out; i++)
    if (pos[i] >= total_len || fabs(err[i]) <= tol)
      return 0;
  double maxerr = err[0];
  for (size_t i = 1; i < n_out; i++)
    if (fabs(maxerr) < fabs(err[i]))
      maxerr = err[i];
  fesetround(FE_TONEAREST);
  feclearexcept(FE_INVALID);
  const long long maxint = llrint(fabs(maxerr)); /* quirk: not divided by tol (:89) */
  if (fetestexcept(FE_INVALID))
    return 0;
  const int width = width_of(maxint);
  uint64_t* mag = (uint64_t*)calloc(total_len, 8);
  uint8_t* sign = (uint8_t*)malloc(total_len);
  memset(sign, 1, total_len);
  const double inv = 1.0 / tol;
  for (size_t i = 0; i < n_out; i++) {
    const long long ll = llrint(err[i] * inv);
    sign[pos[i]] = (ll >= 0);
    uint64_t m = (uint64_t)llabs(ll);
    if (width == 1)
      m = (uint8_t)m;
    else if (width == 2)
      m = (uint16_t)m;
    else if (width == 4)
      m = (uint32_t)m;
    mag[pos[i]] = m;
  }
  size_t len = so_speck1d_encode(mag, sign, total_len, out, cap);
  free(mag);
  free(sign);
  return len;
}

/* decode :133-149, m_inverse_quantize :206-234 */
size_t so_outlier_decode(const uint8_t* stream, size_t len, size_t total_len, double tol,
                         uint64_t* pos, double* err, size_t cap)
{
  uint64_t* mag = (uint64_t*)malloc(total_len * 8);
  uint8_t* sign = (uint8_t*)malloc(total_len);
  so_speck1d_decode(stream, len, total_len, mag, sign);
  size_t k = 0;
  for (size_t i = 0; i < total_len; i++) {
    if (mag[i] == 0)
      continue;
    double e = (mag[i] == 1) ? 1.1 : (double)mag[i] - 0.25;
    e *= (tol * (sign[i] ? 1.0 : -1.0));
    if (k < cap) {
      pos[k] = i;
      err[k] = e;
    }
    k++;
  }
  free(mag);
  free(sign);
  return k;
}

/* ------------------------------------------------------------------------------------------ */
/* one chunk: SPECK_FLT::compress (src/SPECK_FLT.cpp:401-541), decompress (:543-606),         */
/* use_bitstream (:27-109), append_encoded_bitstream (:111-124)                               */
/* ------------------------------------------------------------------------------------------ */

static double estimate_q(int mode, double quality, double param, int high_prec, const double* v,
                         size_t n)
{
  if (mode == 2) { /* PSNR, :271-279 */
    const double t_mse = (param * param) * pow(10.0, -quality / 10.0);
    double q = 2.0 * sqrt(t_mse * 3.0);
    while (so_estimate_mse_midtread(v, n, q) > t_mse)
      q /= exp2(0.25);
    return q;
  }
  if (mode == 3)
    return quality * 1.5;
  if (!high_prec)
    return param / (double)UINT32_MAX;
  return param / 0x1.fffffffffffffp52;
}

size_t so_chunk_compress(double* vals, size_t nx, size_t ny, size_t nz, int mode, double quality,
                         int is_2d, uint8_t** out)
{
  const size_t n = nx * ny * nz;
  uint8_t condi[17];
  *out = NULL;
  if (so_condition(vals, n, condi)) {
    *out = (uint8_t*)malloc(17);
    memcpy(*out, condi, 17);
    return 17;
  }
  double* orig = NULL;
  double param = 0.0;
  if (mode == 3) {
    orig = (double*)malloc(n * 8);
    memcpy(orig, vals, n * 8);
  }
  else if (mode == 2) {
    double mn = vals[0], mx = vals[0];
    for (size_t i = 1; i < n; i++) {
      if (vals[i] < mn)
        mn = vals[i];
      if (vals[i] > mx)
        mx = vals[i];
    }
    param = mx - mn;
  }
  if (is_2d)
    so_dwt2d(vals, nx, ny);
  else
    so_dwt3d(vals, nx, ny, nz);
  if (mode == 1) {
    double m = vals[0];
    for (size_t i = 1; i < n; i++)
      if (fabs(m) < fabs(vals[i]))
        m = vals[i];
    param = fabs(m);
  }
  uint64_t* mag = (uint64_t*)malloc(n * 8);
  uint8_t* sign = (uint8_t*)malloc(n);
  uint8_t* speck = (uint8_t*)malloc(n * 10 + 64);
  uint8_t* outl = NULL;
  size_t speck_len = 0, outl_len = 0;
  int high_prec = 0;
  for (;;) {
    const double q = estimate_q(mode, quality, param, high_prec, vals, n);
    memcpy(condi + 9, &q, 8);
    int width;
    if (so_quantize(vals, n, q, mag, sign, &width) != 0) {
      free(mag); free(sign); free(speck); free(orig);
      return 0;
    }
    if (mode == 3) {
      double* rec = (double*)malloc(n * 8);
      so_inv_quantize(mag, sign, n, q, rec);
      if (is_2d)
        so_idwt2d(rec, nx, ny);
      else
        so_idwt3d(rec, nx, ny, nz);
      size_t cnt = 0;
      for (size_t i = 0; i < n; i++)
        if (fabs(orig[i] - rec[i]) > quality)
          cnt++;
      if (cnt) {
        uint64_t* pos = (uint64_t*)malloc(cnt * 8);
        double* err = (double*)malloc(cnt * 8);
        size_t k = 0;
        for (size_t i = 0; i < n; i++) {
          const double d = orig[i] - rec[i];
          if (fabs(d) > quality) {
            pos[k] = i;
            err[k++] = d;
          }
        }
        outl = (uint8_t*)malloc(n / 4 + cnt * 16 + 64);
        outl_len = so_outlier_encode(pos, err, cnt, n, quality, outl, n / 4 + cnt * 16 + 64);
        free(pos);
        free(err);
      }
      free(rec);
    }
    size_t budget = 0;
    if (mode == 1)
      budget = (size_t)(quality * (double)n);
    if (is_2d)
      speck_len = so_speck2d_encode(mag, sign, nx, ny, budget, speck, n * 10 + 64);
    else
      speck_len = so_speck3d_encode(mag, sign, nx, ny, nz, budget, speck, n * 10 + 64);
    if (mode == 1 && !high_prec && speck_len * 8 < budget) {
      high_prec = 1;
      continue;
    }
    break;
  }
  const size_t total = 17 + speck_len + outl_len;
  *out = (uint8_t*)malloc(total);
  memcpy(*out, condi, 17);
  memcpy(*out + 17, speck, speck_len);
  if (outl_len)
    memcpy(*out + 17 + speck_len, outl, outl_len);
  free(mag); free(sign); free(speck); free(outl); free(orig);
  return total;
}

int so_chunk_decompress(const uint8_t* p, size_t len, size_t nx, size_t ny, size_t nz, int is_2d,
                        double* out)
{
  const size_t n = nx * ny * nz;
  if (len < 17)
    return -1;
  if (p[0] & 0x01) {
    if (len != 17)
      return -1;
    uint64_t nval;
    memcpy(&nval, p + 1, 8);
    so_inverse_condition(out, n, p);
    return 0;
  }
  double q;
  memcpy(&q, p + 9, 8);
  const uint8_t* sp = p + 17;
  size_t remaining = len - 17;
  uint64_t nbits;
  memcpy(&nbits, sp + 1, 8);
  size_t full = 9 + (size_t)((nbits + 7) / 8);
  size_t speck_len = full < remaining ? full : remaining;
  uint64_t* mag = (uint64_t*)malloc(n * 8);
  uint8_t* sign = (uint8_t*)malloc(n);
  if (is_2d)
    so_speck2d_decode(sp, speck_len, nx, ny, mag, sign);
  else
    so_speck3d_decode(sp, speck_len, nx, ny, nz, mag, sign);
  so_inv_quantize(mag, sign, n, q, out);
  if (is_2d)
    so_idwt2d(out, nx, ny);
  else
    so_idwt3d(out, nx, ny, nz);
  size_t pos = 17 + speck_len;
  if (pos < len) {
    const uint8_t* op = p + pos;
    remaining = len - pos;
    if (remaining >= 9) {
      memcpy(&nbits, op + 1, 8);
      full = 9 + (size_t)((nbits + 7) / 8);
      if (remaining == full) {
        const double tol = q / 1.5;
        size_t cnt = so_outlier_decode(op, full, n, tol, NULL, NULL, 0);
        uint64_t* opos = (uint64_t*)malloc((cnt + 1) * 8);
        double* oerr = (double*)malloc((cnt + 1) * 8);
        so_outlier_decode(op, full, n, tol, opos, oerr, cnt);
        for (size_t k = 0; k < cnt; k++)
          out[opos[k]] += oerr[k];
        free(opos);
        free(oerr);
      }
    }
  }
  so_inverse_condition(out, n, p);
  free(mag);
  free(sign);
  return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* whole volume: SPERR3D_OMP_C (src/SPERR3D_OMP_C.cpp:62-261), SPERR3D_OMP_D                   */
/* (src/SPERR3D_OMP_D.cpp:23-184), C API (src/SPERR_C_API.cpp)                                 */
/* ------------------------------------------------------------------------------------------ */

int so_comp_3d(const void* src, int is_float, size_t dimx, size_t dimy, size_t dimz, size_t chunk_x,
               size_t chunk_y, size_t chunk_z, int mode, double quality, size_t nthreads,
               void** dst, size_t* dst_len)
{
  if (*dst != NULL)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  size_t cd[3] = {chunk_x, chunk_y, chunk_z};
  const size_t vd[3] = {dimx, dimy, dimz};
  for (int i = 0; i < 3; i++) { /* set_dims_and_chunks :22-29 */
    if (cd[i] < 1)
      cd[i] = 1;
    if (cd[i] > vd[i])
      cd[i] = vd[i];
  }
  const size_t nchunks = so_chunk_volume(dimx, dimy, dimz, cd[0], cd[1], cd[2], NULL, 0);
  size_t* ch = (size_t*)malloc(nchunks * 6 * sizeof(size_t));
  so_chunk_volume(dimx, dimy, dimz, cd[0], cd[1], cd[2], ch, nchunks);
  uint8_t** streams = (uint8_t**)calloc(nchunks, sizeof(uint8_t*));
  size_t* lens = (size_t*)calloc(nchunks, sizeof(size_t));
  int fail = 0;
#ifdef _OPENMP
  int nt = nthreads ? (int)nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(dynamic)
#endif
  for (size_t i = 0; i < nchunks; i++) {
    const size_t* c = ch + i * 6;
    const size_t n = c[1] * c[3] * c[5];
    double* buf = (double*)malloc(n * 8);
    size_t k = 0;
    for (size_t z = c[4]; z < c[4] + c[5]; z++) /* m_gather_chunk :237-261 */
      for (size_t y = c[2]; y < c[2] + c[3]; y++) {
        const size_t start = z * dimx * dimy + y * dimx + c[0];
        for (size_t x = 0; x < c[1]; x++)
          buf[k++] = is_float ? (double)((const float*)src)[start + x]
                              : ((const double*)src)[start + x];
      }
    lens[i] = so_chunk_compress(buf, c[1], c[3], c[5], mode, quality, 0, &streams[i]);
    if (lens[i] == 0)
      fail = 1;
    free(buf);
  }
  (void)nthreads;
  int rtn = 0;
  if (fail)
    rtn = -1;
  else { /* m_generate_header :163-234 */
    const size_t hlen = (nchunks > 1 ? 20 : 14) + 4 * nchunks;
    size_t total = hlen;
    for (size_t i = 0; i < nchunks; i++)
      total += lens[i];
    uint8_t* o = (uint8_t*)malloc(total);
    o[0] = 0; /* SPERR_VERSION_MAJOR */
    o[1] = (uint8_t)(0x40 | (is_float ? 0x20 : 0) | (nchunks > 1 ? 0x10 : 0));
    const uint32_t v3[3] = {(uint32_t)dimx, (uint32_t)dimy, (uint32_t)dimz};
    memcpy(o + 2, v3, 12);
    size_t pos = 14;
    if (nchunks > 1) {
      const uint16_t c3[3] = {(uint16_t)cd[0], (uint16_t)cd[1], (uint16_t)cd[2]};
      memcpy(o + pos, c3, 6);
      pos += 6;
    }
    for (size_t i = 0; i < nchunks; i++) {
      const uint32_t l = (uint32_t)lens[i];
      memcpy(o + pos, &l, 4);
      pos += 4;
    }
    for (size_t i = 0; i < nchunks; i++) {
      memcpy(o + pos, streams[i], lens[i]);
      pos += lens[i];
    }
    *dst = o;
    *dst_len = total;
  }
  for (size_t i = 0; i < nchunks; i++)
    free(streams[i]);
  free(streams);
  free(lens);
  free(ch);
  return rtn;
}

int so_decomp_3d(const void* src, size_t src_len, int output_float, size_t nthreads, size_t* dimx,
                 size_t* dimy, size_t* dimz, void** dst)
{
  if (*dst != NULL)
    return 1;
  const uint8_t* p = (const uint8_t*)src;
  if (src_len < 14)
    return -1;
  if (p[0] != 0) /* version, src/SPERR3D_OMP_D.cpp:33 */
    return -1;
  if (!(p[1] & 0x40))
    return -1;
  const int multi = (p[1] & 0x10) != 0;
  uint32_t v3[3];
  memcpy(v3, p + 2, 12);
  size_t cd[3] = {v3[0], v3[1], v3[2]};
  size_t pos = 14;
  if (multi) {
    uint16_t c3[3];
    memcpy(c3, p + 14, 6);
    cd[0] = c3[0]; cd[1] = c3[1]; cd[2] = c3[2];
    pos = 20;
  }
  if (v3[0] == 0 || v3[1] == 0 || v3[2] == 0 || cd[0] == 0 || cd[1] == 0 || cd[2] == 0)
    return -1;
  const size_t nchunks = so_chunk_volume(v3[0], v3[1], v3[2], cd[0], cd[1], cd[2], NULL, 0);
  size_t* ch = (size_t*)malloc(nchunks * 6 * sizeof(size_t));
  so_chunk_volume(v3[0], v3[1], v3[2], cd[0], cd[1], cd[2], ch, nchunks);
  size_t* offs = (size_t*)malloc((nchunks + 1) * sizeof(size_t));
  offs[0] = pos + 4 * nchunks;
  for (size_t i = 0; i < nchunks; i++) {
    uint32_t l;
    memcpy(&l, p + pos + 4 * i, 4);
    offs[i + 1] = offs[i] + l;
  }
  if (offs[nchunks] != src_len) {
    free(ch); free(offs);
    return -1;
  }
  const size_t total = (size_t)v3[0] * v3[1] * v3[2];
  double* vol = (double*)malloc(total * 8);
  int fail = 0;
#ifdef _OPENMP
  int nt = nthreads ? (int)nthreads : omp_get_max_threads();
#pragma omp parallel for num_threads(nt) schedule(dynamic)
#endif
  for (size_t i = 0; i < nchunks; i++) {
    const size_t* c = ch + i * 6;
    const size_t n = c[1] * c[3] * c[5];
    double* buf = (double*)malloc(n * 8);
    if (so_chunk_decompress(p + offs[i], offs[i + 1] - offs[i], c[1], c[3], c[5], 0, buf) != 0)
      fail = 1;
    size_t k = 0;
    for (size_t z = c[4]; z < c[4] + c[5]; z++) /* m_scatter_chunk :167-184 */
      for (size_t y = c[2]; y < c[2] + c[3]; y++) {
        const size_t start = z * (size_t)v3[0] * v3[1] + y * (size_t)v3[0] + c[0];
        for (size_t x = 0; x < c[1]; x++)
          vol[start + x] = buf[k++];
      }
    free(buf);
  }
  (void)nthreads;
  free(ch);
  free(offs);
  if (fail) {
    free(vol);
    return -1;
  }
  *dimx = v3[0]; *dimy = v3[1]; *dimz = v3[2];
  if (output_float) {
    float* f = (float*)malloc(total * 4);
    for (size_t i = 0; i < total; i++)
      f[i] = (float)vol[i];
    free(vol);
    *dst = f;
  }
  else
    *dst = vol;
  return 0;
}

/* src/SPERR_C_API.cpp:11-96 */
int so_comp_2d(const void* src, int is_float, size_t dimx, size_t dimy, int mode, double quality,
               int out_inc_header, void** dst, size_t* dst_len)
{
  if (*dst != NULL)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  const size_t n = dimx * dimy;
  double* buf = (double*)malloc(n * 8);
  for (size_t i = 0; i < n; i++)
    buf[i] = is_float ? (double)((const float*)src)[i] : ((const double*)src)[i];
  uint8_t* s = NULL;
  const size_t len = so_chunk_compress(buf, dimx, dimy, 1, mode, quality, 1, &s);
  free(buf);
  if (len == 0)
    return -1;
  const size_t h = out_inc_header ? 10 : 0;
  uint8_t* o = (uint8_t*)malloc(h + len);
  if (h) {
    o[0] = 0;
    o[1] = (uint8_t)(is_float ? 0x20 : 0);
    const uint32_t d2[2] = {(uint32_t)dimx, (uint32_t)dimy};
    memcpy(o + 2, d2, 8);
  }
  memcpy(o + h, s, len);
  free(s);
  *dst = o;
  *dst_len = h + len;
  return 0;
}

/* src/SPERR_C_API.cpp:98-134 */
int so_decomp_2d(const void* src, size_t src_len, int output_float, size_t dimx, size_t dimy,
                 void** dst)
{
  if (*dst != NULL)
    return 1;
  const size_t n = dimx * dimy;
  double* buf = (double*)malloc(n * 8);
  if (so_chunk_decompress((const uint8_t*)src, src_len, dimx, dimy, 1, 1, buf) != 0) {
    free(buf);
    return -1;
  }
  if (output_float) {
    float* f = (float*)malloc(n * 4);
    for (size_t i = 0; i < n; i++)
      f[i] = (float)buf[i];
    free(buf);
    *dst = f;
  }
  else
    *dst = buf;
  return 0;
}

/* src/SPERR_C_API.cpp:260-281 over SPERR3D_Stream_Tools::progressive_truncate and
 * m_progressive_helper (src/SPERR3D_Stream_Tools.cpp:131-226), extract_sections
 * (src/sperr_helper.cpp:401-427) */
int so_trunc_3d(const void* src, size_t src_len, unsigned pct, void** dst, size_t* dst_len)
{
  if (*dst != NULL)
    return 1;
  const uint8_t* p = (const uint8_t*)src;
  if (src_len < 20)
    return -1;
  const int multi = (p[1] & 0x10) != 0;
  uint32_t v3[3];
  memcpy(v3, p + 2, 12);
  size_t cd[3] = {v3[0], v3[1], v3[2]};
  size_t pos = 14;
  if (multi) {
    uint16_t c3[3];
    memcpy(c3, p + 14, 6);
    cd[0] = c3[0]; cd[1] = c3[1]; cd[2] = c3[2];
    pos = 20;
  }
  const size_t nchunks = so_chunk_volume(v3[0], v3[1], v3[2], cd[0], cd[1], cd[2], NULL, 0);
  const size_t hlen = pos + 4 * nchunks;
  if (src_len < hlen)
    return -1;
  size_t* off = (size_t*)malloc(nchunks * sizeof(size_t));
  size_t* len = (size_t*)malloc(nchunks * sizeof(size_t));
  size_t at = hlen, far = 0, total = hlen;
  const int cut = pct != 0 && pct < 100;
  for (size_t i = 0; i < nchunks; i++) {
    uint32_t l;
    memcpy(&l, p + pos + 4 * i, 4);
    off[i] = at;
    at += l;
    size_t keep = l;
    if (cut && keep > 64) {   /* m_progressive_min_chunk_bytes */
      keep = (size_t)((double)pct / 100.0 * (double)keep);
      if (keep < 64)
        keep = 64;
    }
    len[i] = keep;
    if (off[i] + keep > far)
      far = off[i] + keep;
    total += keep;
  }
  if (src_len < far) {
    free(off);
    free(len);
    return -1;
  }
  uint8_t* o = (uint8_t*)malloc(total);
  memcpy(o, p, hlen);
  if (cut) {
    o[0] = 0;
    o[1] |= 0x80;
    for (size_t i = 0; i < nchunks; i++) {
      const uint32_t l = (uint32_t)len[i];
      memcpy(o + pos + 4 * i, &l, 4);
    }
  }
  size_t w = hlen;
  for (size_t i = 0; i < nchunks; i++) {
    memcpy(o + w, p + off[i], len[i]);
    w += len[i];
  }
  free(off);
  free(len);
  *dst = o;
  *dst_len = total;
  return 0;
}
