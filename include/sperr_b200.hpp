// sperr_b200 -- C++ face of the chunked 3D coder, mirroring the reference's public classes
//   sperr::SPERR3D_OMP_C   /root/reference/include/SPERR3D_OMP_C.h:14-35, src/SPERR3D_OMP_C.cpp:12-161
//   sperr::SPERR3D_OMP_D   /root/reference/include/SPERR3D_OMP_D.h:15-32, src/SPERR3D_OMP_D.cpp:8-135
// (same method names, argument meaning and return codes) so that code written against them -- e.g.
// utilities/sperr3d.cpp:277-329 -- compiles against this header by switching the namespace.
// Header-only: everything goes through the C ABI of libsperr_b200.so (include/sperr_b200.h); the
// chunk loop the reference runs on OpenMP threads runs on the GPU, set_num_threads() is accepted and
// ignored.
#ifndef SPERR_B200_HPP
#define SPERR_B200_HPP

#include <array>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <utility>
#include <vector>

#include "sperr_b200.h"

namespace sperr_b200 {

using dims_type = std::array<size_t, 3>;
using vec8_type = std::vector<uint8_t>;
using vecd_type = std::vector<double>;

// /root/reference/include/sperr_helper.h:54-64
enum class RTNType {
  Good = 0,
  WrongLength,
  IOError,
  BitBudgetMet,
  VersionMismatch,
  SliceVolumeMismatch,
  CompModeUnknown,
  FE_Invalid,
  Error
};

enum class CompMode { Unknown = 0, Rate = 1, PSNR = 2, PWE = 3 };

class SPERR3D_OMP_C {
 public:
  void set_num_threads(size_t) {}   // the chunk loop runs on the GPU

  // chunk_dims is a preferred value, clamped to [1, vol_dims] (src/SPERR3D_OMP_C.cpp:22-29)
  void set_dims_and_chunks(dims_type vol_dims, dims_type chunk_dims)
  {
    m_dims = vol_dims;
    for (size_t i = 0; i < 3; i++) {
      const size_t c = chunk_dims[i] < 1 ? 1 : chunk_dims[i];
      m_chunk_dims[i] = c < vol_dims[i] ? c : vol_dims[i];
    }
  }
  void set_psnr(double v) { m_mode = CompMode::PSNR; m_quality = v; }
  void set_tolerance(double v) { m_mode = CompMode::PWE; m_quality = v; }
  void set_bitrate(double v) { m_mode = CompMode::Rate; m_quality = v; }

  // src/SPERR3D_OMP_C.cpp:62-141
  template <typename T>
  RTNType compress(const T* buf, size_t buf_len)
  {
    static_assert(sizeof(T) == 4 || sizeof(T) == 8, "float or double input");
    if (m_mode == CompMode::Unknown)
      return RTNType::CompModeUnknown;
    if (buf_len != m_dims[0] * m_dims[1] * m_dims[2])
      return RTNType::WrongLength;
    void* out = nullptr;
    size_t len = 0;
    const int rc = sperr_comp_3d(buf, sizeof(T) == 4, m_dims[0], m_dims[1], m_dims[2], m_chunk_dims[0],
                                 m_chunk_dims[1], m_chunk_dims[2], int(m_mode), m_quality, 0, &out, &len);
    if (rc != 0)
      return RTNType::Error;
    m_stream.assign(static_cast<uint8_t*>(out), static_cast<uint8_t*>(out) + len);
    std::free(out);
    return RTNType::Good;
  }

  vec8_type get_encoded_bitstream() const { return m_stream; }

 private:
  CompMode m_mode = CompMode::Unknown;
  double m_quality = 0.0;
  dims_type m_dims = {0, 0, 0}, m_chunk_dims = {0, 0, 0};
  vec8_type m_stream;
};

class SPERR3D_OMP_D {
 public:
  void set_num_threads(size_t) {}

  // Parses the header and keeps the pointer (src/SPERR3D_OMP_D.cpp:23-49).
  RTNType use_bitstream(const void* p, size_t total_len)
  {
    m_ptr = nullptr;
    const uint8_t* u = static_cast<const uint8_t*>(p);
    if (u == nullptr || total_len < 14)
      return RTNType::WrongLength;
    if (u[0] != 0)   // SPERR_VERSION_MAJOR
      return RTNType::VersionMismatch;
    if (!(u[1] & 0x40))
      return RTNType::SliceVolumeMismatch;
    size_t vol[3], chunk[3], hlen = 0, nchunks = 0;
    int is_float = 0;
    if (sperr_b200_parse_container(p, total_len, vol, chunk, &is_float, &hlen, nullptr, 0, &nchunks) != 0)
      return RTNType::WrongLength;
    std::vector<uint32_t> lens(nchunks);
    if (sperr_b200_parse_container(p, total_len, vol, chunk, &is_float, &hlen, lens.data(), nchunks,
                                   &nchunks) != 0)
      return RTNType::WrongLength;
    size_t stream_len = hlen;
    for (uint32_t l : lens)
      stream_len += l;
    if (stream_len != total_len)   // header.stream_len != total_len, src/SPERR3D_OMP_D.cpp:37-38
      return RTNType::WrongLength;
    m_dims = {vol[0], vol[1], vol[2]};
    m_chunk_dims = {chunk[0], chunk[1], chunk[2]};
    m_ptr = u;
    m_len = total_len;
    return RTNType::Good;
  }

  // The pointer MUST be the one given to use_bitstream() (src/SPERR3D_OMP_D.cpp:51-56). With
  // multi_res the coarsened volumes are kept as well (view_hierarchy, coarsest first).
  RTNType decompress(const void* bitstream, bool multi_res = false)
  {
    if (bitstream == nullptr || m_ptr == nullptr || bitstream != m_ptr)
      return RTNType::Error;
    void* out = nullptr;
    size_t dx = 0, dy = 0, dz = 0;
    m_hierarchy.clear();
    if (!multi_res) {
      if (sperr_decomp_3d(bitstream, m_len, 0, 0, &dx, &dy, &dz, &out) != 0)
        return RTNType::Error;
    }
    else {
      size_t nlev = 0, ld[24];
      void* lv[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
      if (sperr_b200_decomp_3d_multires(bitstream, m_len, 0, &dx, &dy, &dz, &out, &nlev, ld, lv) != 0)
        return RTNType::Error;
      for (size_t h = 0; h < nlev; h++) {
        const double* p = static_cast<const double*>(lv[h]);
        m_hierarchy.emplace_back(p, p + ld[3 * h] * ld[3 * h + 1] * ld[3 * h + 2]);
        std::free(lv[h]);
      }
    }
    const double* d = static_cast<const double*>(out);
    m_vol.assign(d, d + dx * dy * dz);
    std::free(out);
    return RTNType::Good;
  }

  const vecd_type& view_decoded_data() const { return m_vol; }
  vecd_type&& release_decoded_data() { return std::move(m_vol); }
  const std::vector<vecd_type>& view_hierarchy() const { return m_hierarchy; }
  std::vector<vecd_type>&& release_hierarchy() { return std::move(m_hierarchy); }
  dims_type get_dims() const { return m_dims; }
  dims_type get_chunk_dims() const { return m_chunk_dims; }

 private:
  dims_type m_dims = {0, 0, 0}, m_chunk_dims = {0, 0, 0};
  const uint8_t* m_ptr = nullptr;
  size_t m_len = 0;
  vecd_type m_vol;
  std::vector<vecd_type> m_hierarchy;
};

}  // namespace sperr_b200

#endif
