// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Stage-level C entry points over the UNMODIFIED reference classes, so that the
// oracle restatement (oracle/sperr_oracle.c) and the CUDA kernels can be checked
// stage by stage against the real reference, not only end to end.
//
// This file is ours; it is compiled together with the reference's own sources
// (taken where they lie under /root/reference, never copied) into
// oracle/_ref/libsperr_ref.so by oracle/Makefile. The reference's C API
// (sperr_comp_3d, sperr_decomp_3d, ...) is exported by the same library.
//
// Reference classes wrapped (file:line in /root/reference):
//   Conditioner::condition / inverse_condition   src/Conditioner.cpp:10-96
//   CDF97::dwt3d / idwt3d / dwt2d / idwt2d       src/CDF97.cpp:102-148
//   SPECK3D_INT_ENC / _DEC                       src/SPECK3D_INT_ENC.cpp, src/SPECK_INT.cpp:110-228
//   SPECK2D_INT_ENC / _DEC                       src/SPECK2D_INT_ENC.cpp
//   Outlier_Coder::encode / decode               src/Outlier_Coder.cpp:71-149
//   chunk_volume, num_of_xforms, ...             src/sperr_helper.cpp

#include <cstdlib>
#include <cstring>
#include <vector>

#include "CDF97.h"
#include "Conditioner.h"
#include "Outlier_Coder.h"
#include "SPECK2D_INT_DEC.h"
#include "SPECK2D_INT_ENC.h"
#include "SPECK3D_INT_DEC.h"
#include "SPECK3D_INT_ENC.h"
#include "sperr_helper.h"

#ifdef USE_OMP
#include <omp.h>
#endif

namespace {

template <typename ENC, typename T>
size_t int_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny, size_t nz,
                  size_t budget_bits, uint8_t* out, size_t out_cap)
{
  const size_t n = nx * ny * nz;
  std::vector<T> coeffs(n);
  for (size_t i = 0; i < n; i++)
    coeffs[i] = static_cast<T>(mags[i]);
  sperr::Bitmask mask(n);
  for (size_t i = 0; i < n; i++)
    mask.wbit(i, signs[i] != 0);
  ENC enc;
  enc.set_dims({nx, ny, nz});
  enc.set_budget(budget_bits);
  enc.use_coeffs(std::move(coeffs), std::move(mask));
  enc.encode();
  sperr::vec8_type buf;
  enc.append_encoded_bitstream(buf);
  if (buf.size() <= out_cap)
    std::memcpy(out, buf.data(), buf.size());
  return buf.size();
}

template <typename DEC>
void int_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, size_t nz, uint64_t* mags,
                uint8_t* signs)
{
  DEC dec;
  dec.set_dims({nx, ny, nz});
  dec.use_bitstream(stream, len);
  dec.decode();
  const auto& c = dec.view_coeffs();
  const auto& s = dec.view_signs();
  for (size_t i = 0; i < c.size(); i++) {
    mags[i] = c[i];
    signs[i] = s.rbit(i);
  }
}

}  // namespace

extern "C" {

int ref_omp_max_threads()
{
#ifdef USE_OMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Conditioner: in-place on `buf`, writes the 17-byte header.
void ref_condition(double* buf, size_t nx, size_t ny, size_t nz, uint8_t* header17)
{
  const size_t n = nx * ny * nz;
  sperr::vecd_type v(buf, buf + n);
  sperr::Conditioner c;
  auto h = c.condition(v, {nx, ny, nz});
  std::memcpy(header17, h.data(), 17);
  std::memcpy(buf, v.data(), n * sizeof(double));
}

void ref_dwt3d(double* buf, size_t nx, size_t ny, size_t nz, int inverse)
{
  const size_t n = nx * ny * nz;
  sperr::CDF97 cdf;
  cdf.copy_data(buf, n, {nx, ny, nz});
  if (inverse)
    cdf.idwt3d();
  else
    cdf.dwt3d();
  std::memcpy(buf, cdf.view_data().data(), n * sizeof(double));
}

void ref_dwt2d(double* buf, size_t nx, size_t ny, int inverse)
{
  const size_t n = nx * ny;
  sperr::CDF97 cdf;
  cdf.copy_data(buf, n, {nx, ny, 1});
  if (inverse)
    cdf.idwt2d();
  else
    cdf.dwt2d();
  std::memcpy(buf, cdf.view_data().data(), n * sizeof(double));
}

// Integer SPECK3D. `width` = 1, 2, 4, 8 selects the reference's integer type.
// Returns the stream length in bytes (the stream is only written if it fits `out_cap`).
size_t ref_speck3d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                          size_t nz, int width, size_t budget_bits, uint8_t* out, size_t out_cap)
{
  using namespace sperr;
  switch (width) {
    case 1:
      return int_encode<SPECK3D_INT_ENC<uint8_t>, uint8_t>(mags, signs, nx, ny, nz, budget_bits, out, out_cap);
    case 2:
      return int_encode<SPECK3D_INT_ENC<uint16_t>, uint16_t>(mags, signs, nx, ny, nz, budget_bits, out, out_cap);
    case 4:
      return int_encode<SPECK3D_INT_ENC<uint32_t>, uint32_t>(mags, signs, nx, ny, nz, budget_bits, out, out_cap);
    default:
      return int_encode<SPECK3D_INT_ENC<uint64_t>, uint64_t>(mags, signs, nx, ny, nz, budget_bits, out, out_cap);
  }
}

void ref_speck3d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, size_t nz,
                        uint64_t* mags, uint8_t* signs)
{
  using namespace sperr;
  const auto planes = speck_int_get_num_bitplanes(stream);
  if (planes <= 8)
    int_decode<SPECK3D_INT_DEC<uint8_t>>(stream, len, nx, ny, nz, mags, signs);
  else if (planes <= 16)
    int_decode<SPECK3D_INT_DEC<uint16_t>>(stream, len, nx, ny, nz, mags, signs);
  else if (planes <= 32)
    int_decode<SPECK3D_INT_DEC<uint32_t>>(stream, len, nx, ny, nz, mags, signs);
  else
    int_decode<SPECK3D_INT_DEC<uint64_t>>(stream, len, nx, ny, nz, mags, signs);
}

size_t ref_speck2d_encode(const uint64_t* mags, const uint8_t* signs, size_t nx, size_t ny,
                          int width, size_t budget_bits, uint8_t* out, size_t out_cap)
{
  using namespace sperr;
  switch (width) {
    case 1:
      return int_encode<SPECK2D_INT_ENC<uint8_t>, uint8_t>(mags, signs, nx, ny, 1, budget_bits, out, out_cap);
    case 2:
      return int_encode<SPECK2D_INT_ENC<uint16_t>, uint16_t>(mags, signs, nx, ny, 1, budget_bits, out, out_cap);
    case 4:
      return int_encode<SPECK2D_INT_ENC<uint32_t>, uint32_t>(mags, signs, nx, ny, 1, budget_bits, out, out_cap);
    default:
      return int_encode<SPECK2D_INT_ENC<uint64_t>, uint64_t>(mags, signs, nx, ny, 1, budget_bits, out, out_cap);
  }
}

void ref_speck2d_decode(const uint8_t* stream, size_t len, size_t nx, size_t ny, uint64_t* mags,
                        uint8_t* signs)
{
  using namespace sperr;
  const auto planes = speck_int_get_num_bitplanes(stream);
  if (planes <= 8)
    int_decode<SPECK2D_INT_DEC<uint8_t>>(stream, len, nx, ny, 1, mags, signs);
  else if (planes <= 16)
    int_decode<SPECK2D_INT_DEC<uint16_t>>(stream, len, nx, ny, 1, mags, signs);
  else if (planes <= 32)
    int_decode<SPECK2D_INT_DEC<uint32_t>>(stream, len, nx, ny, 1, mags, signs);
  else
    int_decode<SPECK2D_INT_DEC<uint64_t>>(stream, len, nx, ny, 1, mags, signs);
}

// Outlier coder. Returns the stream length (0 on error).
size_t ref_outlier_encode(const uint64_t* pos, const double* err, size_t n_out, size_t total_len,
                          double tol, uint8_t* out, size_t out_cap)
{
  sperr::Outlier_Coder oc;
  oc.set_length(total_len);
  oc.set_tolerance(tol);
  std::vector<sperr::Outlier> los;
  los.reserve(n_out);
  for (size_t i = 0; i < n_out; i++)
    los.emplace_back(pos[i], err[i]);
  oc.use_outlier_list(std::move(los));
  if (oc.encode() != sperr::RTNType::Good)
    return 0;
  sperr::vec8_type buf;
  oc.append_encoded_bitstream(buf);
  if (buf.size() <= out_cap)
    std::memcpy(out, buf.data(), buf.size());
  return buf.size();
}

// Returns the number of outliers recovered (written up to `cap`).
size_t ref_outlier_decode(const uint8_t* stream, size_t len, size_t total_len, double tol,
                          uint64_t* pos, double* err, size_t cap)
{
  sperr::Outlier_Coder oc;
  oc.set_length(total_len);
  oc.set_tolerance(tol);
  if (oc.use_bitstream(stream, len) != sperr::RTNType::Good)
    return 0;
  if (oc.decode() != sperr::RTNType::Good)
    return 0;
  const auto& los = oc.view_outlier_list();
  for (size_t i = 0; i < los.size() && i < cap; i++) {
    pos[i] = los[i].pos;
    err[i] = los[i].err;
  }
  return los.size();
}

// Geometry helpers.
size_t ref_num_of_xforms(size_t len) { return sperr::num_of_xforms(len); }
size_t ref_num_of_partitions(size_t len) { return sperr::num_of_partitions(len); }
int ref_can_use_dyadic(size_t nx, size_t ny, size_t nz)
{
  auto d = sperr::can_use_dyadic({nx, ny, nz});
  return d ? int(*d) : -1;
}
void ref_calc_approx_detail_len(size_t len, size_t lev, size_t* out2)
{
  auto a = sperr::calc_approx_detail_len(len, lev);
  out2[0] = a[0];
  out2[1] = a[1];
}
// Writes up to cap chunks (6 size_t each); returns the number of chunks.
size_t ref_chunk_volume(size_t vx, size_t vy, size_t vz, size_t cx, size_t cy, size_t cz,
                        size_t* out, size_t cap)
{
  auto c = sperr::chunk_volume({vx, vy, vz}, {cx, cy, cz});
  for (size_t i = 0; i < c.size() && i < cap; i++)
    for (size_t j = 0; j < 6; j++)
      out[i * 6 + j] = c[i][j];
  return c.size();
}

}  // extern "C"

// ---- multi-resolution decoding through the reference's own class (SPERR3D_OMP_D.cpp:51-135) ----
#include "SPERR3D_OMP_D.h"

extern "C" int ref_decomp_3d_multires(const void* src, size_t len, size_t* nlevels, size_t* level_dims,
                                      double** levels, double** full, size_t* dims3)
{
  sperr::SPERR3D_OMP_D dec;
  dec.set_num_threads(1);   // one decompressor object: a constant chunk then shows the reference's stale-data quirk deterministically
  if (dec.use_bitstream(src, len) != sperr::RTNType::Good)
    return -1;
  if (dec.decompress(src, true) != sperr::RTNType::Good)
    return -1;
  const auto d = dec.get_dims();
  for (int i = 0; i < 3; i++)
    dims3[i] = d[i];
  const auto& vol = dec.view_decoded_data();
  *full = static_cast<double*>(std::malloc(vol.size() * sizeof(double)));
  std::memcpy(*full, vol.data(), vol.size() * sizeof(double));
  const auto& h = dec.view_hierarchy();
  const auto res = sperr::coarsened_resolutions(d, dec.get_chunk_dims());
  *nlevels = h.size();
  for (size_t i = 0; i < h.size() && i < 8; i++) {
    for (int k = 0; k < 3; k++)
      level_dims[3 * i + k] = res[i][k];
    levels[i] = static_cast<double*>(std::malloc(h[i].size() * sizeof(double)));
    std::memcpy(levels[i], h[i].data(), h[i].size() * sizeof(double));
  }
  return 0;
}

// ---- SPERR3D_Stream_Tools through the reference's own class (src/SPERR3D_Stream_Tools.cpp) ----
#include "SPERR3D_Stream_Tools.h"

extern "C" size_t ref_tools_header_len(const uint8_t* magic20)
{
  std::array<uint8_t, 20> a{};
  std::memcpy(a.data(), magic20, 20);
  return sperr::SPERR3D_Stream_Tools().get_header_len(a);
}

// fields[0..11] = major, is_portion, is_3D, is_float, multi_chunk, vol xyz, chunk xyz, header_len;
// fields[12] = stream_len; offsets receives up to cap {offset, length} entries; returns their count
extern "C" size_t ref_tools_stream_header(const void* p, size_t* fields, size_t* offsets, size_t cap)
{
  const auto h = sperr::SPERR3D_Stream_Tools().get_stream_header(p);
  const size_t f[13] = {h.major_version, h.is_portion, h.is_3D, h.is_float, h.multi_chunk,
                        h.vol_dims[0], h.vol_dims[1], h.vol_dims[2], h.chunk_dims[0], h.chunk_dims[1],
                        h.chunk_dims[2], h.header_len, h.stream_len};
  std::memcpy(fields, f, sizeof(f));
  for (size_t i = 0; i < h.chunk_offsets.size() && i < cap; i++)
    offsets[i] = h.chunk_offsets[i];
  return h.chunk_offsets.size();
}

extern "C" size_t ref_tools_progressive_read(const char* filename, unsigned pct, uint8_t* out, size_t cap)
{
  const auto v = sperr::SPERR3D_Stream_Tools().progressive_read(filename, pct);
  if (v.size() <= cap)
    std::memcpy(out, v.data(), v.size());
  return v.size();
}
