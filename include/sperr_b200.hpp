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
#include <cstdio>
#include <cstring>
#include <string>
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

// ---------------------------------------------------------------------------------------------
// sperr::SPERR3D_Header / sperr::SPERR3D_Stream_Tools
//   /root/reference/include/SPERR3D_Stream_Tools.h:11-46, src/SPERR3D_Stream_Tools.cpp:11-226
// Host-only container tools: header parsing and progressive access (keep pct % of every chunk,
// at least 64 bytes, header flagged is_portion). Same method names and results; the geometry and
// the truncation itself go through the C ABI (sperr_b200_parse_container, sperr_trunc_3d).
// ---------------------------------------------------------------------------------------------
struct SPERR3D_Header {
  uint8_t major_version = 0;
  bool is_portion = false;
  bool is_3D = false;
  bool is_float = false;
  bool multi_chunk = false;
  dims_type vol_dims = {0, 0, 0};
  dims_type chunk_dims = {0, 0, 0};
  size_t header_len = 0;
  size_t stream_len = 0;
  std::vector<size_t> chunk_offsets;   // {offset, length} pairs, one per chunk
};

class SPERR3D_Stream_Tools {
 public:
  // Total header length from the first 20 bytes of a container (0: not a valid geometry).
  size_t get_header_len(std::array<uint8_t, 20> magic) const
  {
    dims_type v, c;
    const bool multi = m_geometry(magic.data(), v, c);
    size_t vol[3] = {v[0], v[1], v[2]}, chunk[3] = {c[0], c[1], c[2]};
    const size_t n = sperr_b200_num_chunks(vol, chunk);
    return n == 0 ? 0 : (multi ? 20 : 14) + 4 * n;
  }

  // `p` must hold at least get_header_len() bytes.
  SPERR3D_Header get_stream_header(const void* p) const
  {
    SPERR3D_Header h;
    const uint8_t* u = static_cast<const uint8_t*>(p);
    h.major_version = u[0];
    h.is_portion = (u[1] & 0x80) != 0;
    h.is_3D = (u[1] & 0x40) != 0;
    h.is_float = (u[1] & 0x20) != 0;
    h.multi_chunk = m_geometry(u, h.vol_dims, h.chunk_dims);
    size_t vol[3] = {h.vol_dims[0], h.vol_dims[1], h.vol_dims[2]};
    size_t chunk[3] = {h.chunk_dims[0], h.chunk_dims[1], h.chunk_dims[2]};
    const size_t n = sperr_b200_num_chunks(vol, chunk);
    const size_t pos = h.multi_chunk ? 20 : 14;
    h.header_len = pos + 4 * n;
    h.stream_len = h.header_len;
    h.chunk_offsets.resize(2 * n);
    for (size_t i = 0; i < n; i++) {
      uint32_t l;
      std::memcpy(&l, u + pos + 4 * i, 4);
      h.chunk_offsets[2 * i] = h.stream_len;
      h.chunk_offsets[2 * i + 1] = l;
      h.stream_len += l;
    }
    return h;
  }

  // Reads only the header and the leading pct % of every chunk of a container file. Empty on any
  // I/O error (src/SPERR3D_Stream_Tools.cpp:107-129).
  vec8_type progressive_read(const std::string& filename, unsigned pct) const
  {
    std::FILE* f = std::fopen(filename.c_str(), "rb");
    if (!f)
      return {};
    vec8_type out;
    std::array<uint8_t, 20> magic{};
    size_t hlen = 0;
    if (std::fread(magic.data(), 1, 20, f) == 20 && (hlen = get_header_len(magic)) >= 14) {
      vec8_type header(hlen);
      std::rewind(f);
      if (std::fread(header.data(), 1, hlen, f) == hlen) {
        const SPERR3D_Header h = get_stream_header(header.data());
        const auto keep = m_kept_lengths(h, pct);
        out = m_new_header(header, h, keep, pct);
        bool ok = true;
        for (size_t i = 0; i < keep.size() && ok; i++) {
          const size_t at = out.size();
          out.resize(at + keep[i]);
          ok = std::fseek(f, long(h.chunk_offsets[2 * i]), SEEK_SET) == 0 &&
               std::fread(out.data() + at, 1, keep[i], f) == keep[i];
        }
        if (!ok)
          out.clear();
      }
    }
    std::fclose(f);
    return out;
  }

  // The same on a container in memory; stream_len only needs to cover what is kept. Empty on
  // error (src/SPERR3D_Stream_Tools.cpp:131-226).
  vec8_type progressive_truncate(const void* stream, size_t stream_len, unsigned pct) const
  {
    void* out = nullptr;
    size_t n = 0;
    if (stream == nullptr || sperr_trunc_3d(stream, stream_len, pct, &out, &n) != 0)
      return {};
    vec8_type v(static_cast<uint8_t*>(out), static_cast<uint8_t*>(out) + n);
    std::free(out);
    return v;
  }

 private:
  static constexpr size_t m_progressive_min_chunk_bytes = 64;

  static bool m_geometry(const uint8_t* u, dims_type& vol, dims_type& chunk)
  {
    const bool multi = (u[1] & 0x10) != 0;
    uint32_t v3[3];
    std::memcpy(v3, u + 2, 12);
    vol = {v3[0], v3[1], v3[2]};
    chunk = vol;
    if (multi) {
      uint16_t c3[3];
      std::memcpy(c3, u + 14, 6);
      chunk = {c3[0], c3[1], c3[2]};
    }
    return multi;
  }
  static std::vector<size_t> m_kept_lengths(const SPERR3D_Header& h, unsigned pct)
  {
    std::vector<size_t> keep(h.chunk_offsets.size() / 2);
    for (size_t i = 0; i < keep.size(); i++) {
      size_t l = h.chunk_offsets[2 * i + 1];
      if (pct != 0 && pct < 100 && l > m_progressive_min_chunk_bytes) {
        const size_t want = size_t(double(pct) / 100.0 * double(l));
        l = want < m_progressive_min_chunk_bytes ? m_progressive_min_chunk_bytes : want;
      }
      keep[i] = l;
    }
    return keep;
  }
  static vec8_type m_new_header(const vec8_type& old, const SPERR3D_Header& h, const std::vector<size_t>& keep,
                                unsigned pct)
  {
    vec8_type out(old.begin(), old.begin() + long(h.header_len));
    if (pct == 0 || pct >= 100)
      return out;   // the complete container, header untouched
    out[0] = 0;       // SPERR_VERSION_MAJOR
    out[1] |= 0x80;   // is_portion
    const size_t pos = h.multi_chunk ? 20 : 14;
    for (size_t i = 0; i < keep.size(); i++) {
      const uint32_t l = uint32_t(keep[i]);
      std::memcpy(out.data() + pos + 4 * i, &l, 4);
    }
    return out;
  }
};

}  // namespace sperr_b200

#endif
