// Test driver for include/sperr_b200.hpp: the reference's sperr3d utility flow
// (/root/reference/utilities/sperr3d.cpp:277-329) written against the class mirrors.
//   classes_main hostonly
//   classes_main run <in.f32> nx ny nz cx cy cz mode quality <out.stream> <out.f64>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iterator>
#include <string>

#include "sperr_b200.hpp"

using namespace sperr_b200;

static int fail(const char* what)
{
  std::fprintf(stderr, "FAIL: %s\n", what);
  return 1;
}

int main(int argc, char** argv)
{
  if (argc >= 2 && std::string(argv[1]) == "hostonly") {
    SPERR3D_OMP_C enc;
    std::vector<float> v(16 * 16 * 16, 1.0f);
    enc.set_dims_and_chunks({16, 16, 16}, {99, 0, 8});
    if (enc.compress(v.data(), v.size()) != RTNType::CompModeUnknown) return fail("mode unknown");
    enc.set_tolerance(1e-3);
    if (enc.compress(v.data(), v.size() - 1) != RTNType::WrongLength) return fail("wrong length");
    SPERR3D_OMP_D dec;
    uint8_t hdr[14 + 4] = {0, 0x40 | 0x20, 16, 0, 0, 0, 16, 0, 0, 0, 16, 0, 0, 0, 17, 0, 0, 0};
    if (dec.decompress(hdr) != RTNType::Error) return fail("decompress before use_bitstream");
    uint8_t bad = hdr[0];
    hdr[0] = 9;
    if (dec.use_bitstream(hdr, sizeof(hdr)) != RTNType::VersionMismatch) return fail("version");
    hdr[0] = bad;
    hdr[1] = 0x20;
    if (dec.use_bitstream(hdr, sizeof(hdr)) != RTNType::SliceVolumeMismatch) return fail("2D flag");
    hdr[1] = 0x60;
    if (dec.use_bitstream(hdr, sizeof(hdr)) != RTNType::WrongLength) return fail("length");
    std::vector<uint8_t> full(14 + 4 + 17, 0);
    std::memcpy(full.data(), hdr, sizeof(hdr));
    if (dec.use_bitstream(full.data(), full.size()) != RTNType::Good) return fail("good header");
    if (dec.get_dims()[2] != 16 || dec.get_chunk_dims()[0] != 16) return fail("dims");
    if (dec.decompress(hdr) != RTNType::Error) return fail("other pointer");
    std::puts("hostonly ok");
    return 0;
  }
  if (argc == 5 && std::string(argv[1]) == "tools") {
    // classes_main tools <container file> pct <out prefix>: SPERR3D_Stream_Tools mirror; writes
    // <prefix>.hdr (text: the header fields and {offset, length} pairs), <prefix>.read
    // (progressive_read) and <prefix>.trunc (progressive_truncate of the whole file)
    const std::string path = argv[2], prefix = argv[4];
    const unsigned pct = unsigned(std::atoi(argv[3]));
    std::ifstream in(path, std::ios::binary);
    std::vector<uint8_t> all((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
    if (all.size() < 20) return fail("container too short");
    SPERR3D_Stream_Tools tools;
    std::array<uint8_t, 20> magic{};
    std::memcpy(magic.data(), all.data(), 20);
    const size_t hlen = tools.get_header_len(magic);
    const SPERR3D_Header h = tools.get_stream_header(all.data());
    if (h.header_len != hlen) return fail("header length");
    std::FILE* f = std::fopen((prefix + ".hdr").c_str(), "w");
    std::fprintf(f, "%u %d %d %d %d %zu %zu %zu %zu %zu %zu %zu %zu\n", unsigned(h.major_version), int(h.is_portion),
                 int(h.is_3D), int(h.is_float), int(h.multi_chunk), h.vol_dims[0], h.vol_dims[1], h.vol_dims[2],
                 h.chunk_dims[0], h.chunk_dims[1], h.chunk_dims[2], h.header_len, h.stream_len);
    for (size_t v : h.chunk_offsets)
      std::fprintf(f, "%zu ", v);
    std::fprintf(f, "\n");
    std::fclose(f);
    const auto r = tools.progressive_read(path, pct);
    std::ofstream((prefix + ".read").c_str(), std::ios::binary).write(reinterpret_cast<const char*>(r.data()), long(r.size()));
    const auto t = tools.progressive_truncate(all.data(), all.size(), pct);
    std::ofstream((prefix + ".trunc").c_str(), std::ios::binary).write(reinterpret_cast<const char*>(t.data()), long(t.size()));
    if (!tools.progressive_read(path + ".does_not_exist", pct).empty()) return fail("missing file");
    return 0;
  }
  if (argc != 13 || std::string(argv[1]) != "run")
    return fail("usage");
  const size_t nx = std::atol(argv[3]), ny = std::atol(argv[4]), nz = std::atol(argv[5]);
  const size_t cx = std::atol(argv[6]), cy = std::atol(argv[7]), cz = std::atol(argv[8]);
  const int mode = std::atoi(argv[9]);
  const double q = std::atof(argv[10]);
  std::ifstream in(argv[2], std::ios::binary);
  std::vector<char> raw((std::istreambuf_iterator<char>(in)), std::istreambuf_iterator<char>());
  if (raw.size() != nx * ny * nz * 4) return fail("input size");
  SPERR3D_OMP_C enc;
  enc.set_num_threads(0);
  enc.set_dims_and_chunks({nx, ny, nz}, {cx, cy, cz});
  if (mode == 1) enc.set_bitrate(q);
  else if (mode == 2) enc.set_psnr(q);
  else enc.set_tolerance(q);
  if (enc.compress(reinterpret_cast<const float*>(raw.data()), nx * ny * nz) != RTNType::Good)
    return fail("compress");
  const vec8_type stream = enc.get_encoded_bitstream();
  std::ofstream(argv[11], std::ios::binary).write(reinterpret_cast<const char*>(stream.data()), stream.size());
  SPERR3D_OMP_D dec;
  if (dec.use_bitstream(stream.data(), stream.size()) != RTNType::Good) return fail("use_bitstream");
  if (dec.decompress(stream.data()) != RTNType::Good) return fail("decompress");
  const vecd_type vol = dec.release_decoded_data();
  if (vol.size() != nx * ny * nz) return fail("decoded size");
  std::ofstream(argv[12], std::ios::binary).write(reinterpret_cast<const char*>(vol.data()), vol.size() * 8);
  std::puts("run ok");
  return 0;
}
