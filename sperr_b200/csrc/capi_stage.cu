// Stage-level C entry points (host buffers in, host buffers out). They exist so that the parity
// tests can pin each kernel family against the oracle separately; they run the same kernels as
// the full pipeline. Declared in include/sperr_b200.h.
#include "../../include/sperr_b200.h"

#include "batch.h"
#include "outlier.h"
#include "speck.h"

using namespace sperr_b200;

namespace {

template <typename F>
int guarded(F&& f)
{
  try {
    return f();
  }
  catch (const std::exception& e) {
    if (std::getenv("SPERR_B200_VERBOSE"))
      std::fprintf(stderr, "sperr_b200: %s\n", e.what());
    return -1;
  }
}

Chunk whole(size_t nx, size_t ny, size_t nz)
{
  Chunk c;
  c.x0 = c.y0 = c.z0 = 0;
  c.lx = uint32_t(nx); c.ly = uint32_t(ny); c.lz = uint32_t(nz);
  return c;
}

}  // namespace

extern "C" {

int sperr_b200_stage_condition(const void* src, int is_float, size_t nx, size_t ny, size_t nz,
                               double* out_vals, double* out_mean, int* out_is_const)
{
  return guarded([&] {
    cudaStream_t st = 0;
    BatchBuffers b;
    b.setup({whole(nx, ny, nz)}, true, false, false, st);
    const size_t n = nx * ny * nz;
    rt::DBuf d_src(n * (is_float ? 4 : 8));
    rt::h2d(d_src.p, src, d_src.bytes, st);
    SrcVol sv{d_src.p, is_float, nx, ny};
    const unsigned ns = unsigned(mean_num_strides(n));
    rt::DBuf d_sm(ns * 8), d_ns(4), d_nc(4);
    rt::h2d(d_ns.p, &ns, 4, st);
    rt::dset(d_nc.p, 0, 4, st);
    launch_stats(sv, b.dev(), 1, d_sm.as<double>(), int(ns), d_ns.as<unsigned>(), d_nc.as<unsigned>(),
                 false, st);
    launch_gather(sv, b.dev(), 1, n, st);
    b.pull(st);
    *out_mean = b.h[0].mean;
    *out_is_const = b.h[0].is_const;
    if (!b.h[0].is_const)
      rt::d2h(out_vals, b.h[0].coef, n * 8, st);
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_stage_dwt(double* buf, size_t nx, size_t ny, size_t nz, int inverse, int is_2d)
{
  return guarded([&] {
    cudaStream_t st = 0;
    BatchBuffers b;
    b.setup({whole(nx, ny, nz)}, true, false, false, st);
    const size_t n = nx * ny * nz;
    rt::h2d(b.h[0].coef, buf, n * 8, st);
    rt::DBuf ids(4);
    const int zero = 0;
    rt::h2d(ids.p, &zero, 4, st);
    launch_dwt(inverse != 0, b.dev(), ids.as<int>(), 1, uint32_t(nx), uint32_t(ny), uint32_t(nz),
               is_2d != 0, st);
    rt::d2h(buf, b.h[0].coef, n * 8, st);
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_stage_dwt_fused(double* buf, size_t nx, size_t ny, size_t nz, int inverse)
{
  return guarded([&] {
    if (can_use_dyadic(nx, ny, nz) < 1)
      return -2;   // not a shape the fused kernels handle
    cudaStream_t st = 0;
    BatchBuffers b;
    b.setup({whole(nx, ny, nz)}, true, false, false, st);
    if (!b.h[0].fused)
      return -2;
    const size_t n = nx * ny * nz;
    rt::DBuf vol(n * 8), ids(4);
    const int zero = 0;
    rt::h2d(ids.p, &zero, 4, st);
    SrcVol sv{vol.p, 0, nx, ny};
    if (!inverse) {
      rt::h2d(vol.p, buf, n * 8, st);   // mean is 0: v - 0.0 == v bit for bit
      launch_dwt_fused_forward(sv, b.dev(), ids.as<int>(), 1, uint32_t(nx), uint32_t(ny), uint32_t(nz), st);
      rt::d2h(buf, b.h[0].coef, n * 8, st);
    }
    else {
      rt::h2d(b.h[0].coef, buf, n * 8, st);
      launch_dwt_fused_inverse(sv, 0, b.dev(), ids.as<int>(), 1, uint32_t(nx), uint32_t(ny), uint32_t(nz), 0.0,
                               OutlierSink{}, CorrectorList{}, st);
      rt::d2h(buf, vol.p, n * 8, st);
    }
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_stage_quantize(const double* vals, size_t nx, size_t ny, size_t nz, double q,
                              uint64_t* mags, uint8_t* signs, int* wide)
{
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t n = nx * ny * nz;
    BatchBuffers b;
    b.setup({whole(nx, ny, nz)}, true, true, false, st);
    rt::h2d(b.h[0].coef, vals, n * 8, st);
    b.h[0].q = q;
    b.push(st);
    launch_absmax(b.dev(), 1, n, st);
    launch_qdecide(b.dev(), 1, st);
    b.pull(st);
    if (b.h[0].fe_invalid)
      return -1;
    if (b.h[0].wide) {
      b.make_wide(st);
      b.push(st);
    }
    launch_quantize(b.dev(), 1, n, st);
    std::vector<uint32_t> sw((n + 31) / 32);
    rt::d2h(sw.data(), b.h[0].signs, sw.size() * 4, st);
    if (b.h[0].wide)
      rt::d2h(mags, b.h[0].mag, n * 8, st);
    else {
      std::vector<uint32_t> m(n);
      rt::d2h(m.data(), b.h[0].mag, n * 4, st);
      rt::sync(st);
      for (size_t i = 0; i < n; i++)
        mags[i] = m[i];
    }
    rt::sync(st);
    for (size_t i = 0; i < n; i++)
      signs[i] = (sw[i >> 5] >> (i & 31)) & 1u;
    *wide = b.h[0].wide;
    return 0;
  });
}

static int stage_speck_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                              size_t nz, bool is_2d, size_t budget_bits, uint8_t* out, size_t cap,
                              size_t* out_len)
{
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t n = nx * ny * nz;
    bool wide = false;
    for (size_t i = 0; i < n; i++)
      if (mags[i] > 0xFFFFFFFFull)
        wide = true;
    BatchBuffers b;
    b.setup({whole(nx, ny, nz)}, false, true, wide, st, true, is_2d);
    std::vector<uint32_t> sw((n + 31) / 32, 0);
    std::vector<int8_t> pl(n);
    for (size_t i = 0; i < n; i++) {
      if (signs[i])
        sw[i >> 5] |= 1u << (i & 31);
      int p = -1;
      for (uint64_t v = mags[i]; v; v >>= 1)
        p++;
      pl[i] = int8_t(p);
    }
    if (wide)
      rt::h2d(b.h[0].mag, mags, n * 8, st);
    else {
      std::vector<uint32_t> m(n);
      for (size_t i = 0; i < n; i++)
        m[i] = uint32_t(mags[i]);
      rt::h2d(b.h[0].mag, m.data(), n * 4, st);
      rt::sync(st);
    }
    rt::h2d(b.h[0].signs, sw.data(), sw.size() * 4, st);
    rt::h2d(b.h[0].pleaf, pl.data(), n, st);
    if (budget_bits) {
      while (budget_bits % 8)
        budget_bits++;
      b.h[0].budget = budget_bits;
    }
    b.push(st);
    Speck3DEncoder enc;
    std::vector<EncResult> res;
    enc.encode(b.dev(), b.h, b.dev_shapes(), b.shapes, res, st);
    const size_t len = 9 + res[0].payload_bytes;
    *out_len = len;
    if (len > cap)
      return -2;
    out[0] = uint8_t(res[0].planes);
    const uint64_t tb = res[0].total_bits;
    std::memcpy(out + 1, &tb, 8);
    rt::d2h(out + 9, res[0].payload, res[0].payload_bytes, st);
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_stage_speck3d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                                    size_t nz, size_t budget_bits, uint8_t* out, size_t cap,
                                    size_t* out_len)
{
  return stage_speck_encode(mags, signs, nx, ny, nz, false, budget_bits, out, cap, out_len);
}

int sperr_b200_stage_speck2d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                                    size_t budget_bits, uint8_t* out, size_t cap, size_t* out_len)
{
  return stage_speck_encode(mags, signs, nx, ny, 1, true, budget_bits, out, cap, out_len);
}

int sperr_b200_stage_outlier_encode(const uint64_t* pos, const double* err, size_t n_out,
                                    size_t total_len, double tol, uint8_t* out, size_t cap,
                                    size_t* out_len)
{
  return guarded([&] {
    cudaStream_t st = 0;
    if (n_out == 0 || total_len == 0 || tol <= 0.0)
      return -1;
    std::vector<unsigned> p32(n_out);
    for (size_t i = 0; i < n_out; i++)
      p32[i] = unsigned(pos[i]);
    OutlierCoder oc;
    oc.set_outliers({0ull, (unsigned long long)n_out}, p32.data(), err, st);
    std::vector<EncResult> res;
    oc.encode({(unsigned long long)total_len}, tol, res, st);
    const size_t len = 9 + res[0].payload_bytes;
    *out_len = len;
    if (len > cap)
      return -2;
    out[0] = uint8_t(res[0].planes);
    const uint64_t tb = res[0].total_bits;
    std::memcpy(out + 1, &tb, 8);
    rt::d2h(out + 9, res[0].payload, res[0].payload_bytes, st);
    rt::sync(st);
    return 0;
  });
}

}  // extern "C"
