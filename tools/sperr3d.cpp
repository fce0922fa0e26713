// sperr3d -- command-line front end of the 3D coder, flag-compatible with the reference's utility
// (/root/reference/utilities/sperr3d.cpp:97-421; option names, checks and messages follow it), built
// on the class mirrors of include/sperr_b200.hpp, i.e. on the C ABI of libsperr_b200.so. Host code
// only: argument parsing (CLI11 is not available here, so a small parser of its own), file I/O and
// the quality statistics of --print_stats (calc_stats / calc_mean_var, src/sperr_helper.cpp:429-640).
//
//   sperr3d -c --ftype 32 --dims 128 128 128 --pwe 1e-3 --bitstream out.sperr in.float
//   sperr3d -d --decomp_f out.float out.sperr
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <numeric>
#include <string>

#include "sperr_b200.hpp"

namespace sp = sperr_b200;

namespace {

std::vector<uint8_t> read_whole_file(const std::string& name, bool& ok)
{
  std::vector<uint8_t> buf;
  ok = false;
  std::FILE* f = std::fopen(name.c_str(), "rb");
  if (!f)
    return buf;
  std::fseek(f, 0, SEEK_END);
  const long len = std::ftell(f);
  std::fseek(f, 0, SEEK_SET);
  if (len >= 0) {
    buf.resize(size_t(len));
    ok = std::fread(buf.data(), 1, buf.size(), f) == buf.size();
  }
  std::fclose(f);
  return buf;
}

bool write_bytes(const std::string& name, const void* p, size_t n)
{
  std::FILE* f = std::fopen(name.c_str(), "wb");
  if (!f)
    return false;
  const bool ok = std::fwrite(p, 1, n, f) == n;
  std::fclose(f);
  return ok;
}

// utilities/sperr3d.cpp:56-76
int output_buffer(const sp::vecd_type& buf, const std::string& name_f64, const std::string& name_f32)
{
  if (!name_f64.empty() && !write_bytes(name_f64, buf.data(), buf.size() * 8)) {
    std::cout << "Writing decompressed data failed: " << name_f64 << std::endl;
    return 1;
  }
  if (!name_f32.empty()) {
    std::vector<float> f(buf.size());
    std::copy(buf.cbegin(), buf.cend(), f.begin());
    if (!write_bytes(name_f32, f.data(), f.size() * 4)) {
      std::cout << "Writing decompressed data failed: " << name_f32 << std::endl;
      return 1;
    }
  }
  return 0;
}

// utilities/sperr3d.cpp:15-67: one file per coarsened level, named <name>.<X>x<Y>x<Z>
int output_hierarchy(const std::vector<sp::vecd_type>& hierarchy, const std::vector<std::array<size_t, 3>>& res,
                     const std::string& lowres_f64, const std::string& lowres_f32)
{
  auto fname = [&](const std::string& base, size_t i) {
    return base + "." + std::to_string(res[i][0]) + "x" + std::to_string(res[i][1]) + "x" + std::to_string(res[i][2]);
  };
  for (size_t i = 0; i < hierarchy.size(); i++) {
    if (!lowres_f64.empty() && !write_bytes(fname(lowres_f64, i), hierarchy[i].data(), hierarchy[i].size() * 8)) {
      std::cout << "Writing decompressed hierarchy failed: " << fname(lowres_f64, i) << std::endl;
      return 1;
    }
    if (!lowres_f32.empty()) {
      std::vector<float> f(hierarchy[i].size());
      std::copy(hierarchy[i].cbegin(), hierarchy[i].cend(), f.begin());
      if (!write_bytes(fname(lowres_f32, i), f.data(), f.size() * 4)) {
        std::cout << "Writing decompressed hierarchy failed: " << fname(lowres_f32, i) << std::endl;
        return 1;
      }
    }
  }
  return 0;
}

// sperr::coarsened_resolutions(vdims, cdims), src/sperr_helper.cpp:70-123 (3D)
std::vector<std::array<size_t, 3>> coarsened_resolutions(std::array<size_t, 3> v, std::array<size_t, 3> c)
{
  std::vector<std::array<size_t, 3>> out;
  for (int i = 0; i < 3; i++)
    if (c[i] == 0 || v[i] % c[i] != 0)
      return out;
  if (c[2] <= 1 || c[1] < 2)
    return out;
  auto nxf = [](size_t len) { size_t n = 0; while (len >= 9) { n++; len -= len / 2; } return std::min<size_t>(n, 6); };
  const size_t xy = nxf(std::min(c[0], c[1])), z = nxf(c[2]);
  if (!(xy == z || (xy >= 5 && z >= 5)))
    return out;
  const size_t L = std::min(xy, z);
  for (size_t lev = L; lev > 0; lev--) {
    std::array<size_t, 3> r;
    for (int i = 0; i < 3; i++) {
      size_t low = c[i];
      for (size_t k = 0; k < lev; k++)
        low -= low / 2;
      r[i] = low * (v[i] / c[i]);
    }
    out.push_back(r);
  }
  return out;
}

// src/sperr_helper.cpp:429-513: {rmse, linfty, psnr, min, max}, sums taken per stride of 8192
template <typename T>
std::array<T, 5> calc_stats(const T* a, const T* b, size_t n)
{
  const size_t stride = 8192, ns = n / stride;
  const auto mm = std::minmax_element(a, a + n);
  const T amin = *mm.first, amax = *mm.second;
  if (std::equal(a, a + n, b))
    return {T(0), T(0), std::numeric_limits<T>::infinity(), amin, amax};
  std::vector<T> sums(ns + 1, T(0));
  T linf = 0;
  for (size_t s = 0; s <= ns; s++) {
    const size_t lo = s * stride, hi = s < ns ? lo + stride : n;
    T acc = 0;
    for (size_t i = lo; i < hi; i++) {
      const T d = std::abs(a[i] - b[i]);
      linf = std::max(linf, d);
      acc += d * d;
    }
    sums[s] = acc;
  }
  const T mse = std::accumulate(sums.cbegin(), sums.cend(), T(0)) / T(n);
  const T range_sq = (amax - amin) * (amax - amin);
  return {std::sqrt(mse), linf, std::log10(range_sq / mse) * T(10), amin, amax};
}

// src/sperr_helper.cpp:594-640: {mean, variance}, sums taken per stride of 16384
template <typename T>
std::array<T, 2> calc_mean_var(const T* a, size_t n)
{
  const size_t stride = 16384, ns = n / stride;
  std::vector<T> tmp(ns + 1, T(0));
  for (size_t s = 0; s <= ns; s++) {
    const size_t lo = s * stride, hi = s < ns ? lo + stride : n;
    tmp[s] = std::accumulate(a + lo, a + hi, T(0));
  }
  const T mean = std::accumulate(tmp.cbegin(), tmp.cend(), T(0)) / T(n);
  for (size_t s = 0; s <= ns; s++) {
    const size_t lo = s * stride, hi = s < ns ? lo + stride : n;
    tmp[s] = std::accumulate(a + lo, a + hi, T(0), [mean](T init, T v) { return init + (v - mean) * (v - mean); });
  }
  return {mean, std::accumulate(tmp.cbegin(), tmp.cend(), T(0)) / T(n)};
}

void usage()
{
  std::puts(
      "3D SPERR compression and decompression (B200 build)\n\n"
      "Usage: sperr3d [OPTIONS] filename\n\n"
      "  filename                 A data volume to be compressed, or a bitstream to be decompressed.\n"
      "Execution settings:\n"
      "  -c                       Perform a compression task.\n"
      "  -d                       Perform a decompression task.\n"
      "  --omp N                  Accepted for compatibility (the chunk loop runs on the GPU).\n"
      "Input properties (for compression):\n"
      "  --ftype {32,64}          Input float type in bits.\n"
      "  --dims X Y Z             Dimensions of the input volume (fastest-varying first).\n"
      "Output settings:\n"
      "  --bitstream FILE         Output compressed bitstream.\n"
      "  --decomp_f FILE          Output decompressed volume in f32 precision.\n"
      "  --decomp_d FILE          Output decompressed volume in f64 precision.\n"
      "  --decomp_lowres_f FILE   Output lower resolutions of the decompressed volume in f32 precision.\n"
      "  --decomp_lowres_d FILE   Output lower resolutions of the decompressed volume in f64 precision.\n"
      "  --print_stats            Print statistics measuring the compression quality.\n"
      "Compression settings:\n"
      "  --chunks X Y Z           Preferred chunk size. Default: 256 256 256\n"
      "  --pwe TOL | --psnr DB | --bpp RATE");
}

}  // namespace

int main(int argc, char* argv[])
{
  std::string input_file, bitstream, decomp_f32, decomp_f64, lowres_f32, lowres_f64;
  bool cflag = false, dflag = false, print_stats = false;
  size_t ftype = 0, omp = 0;
  std::array<size_t, 3> dims = {0, 0, 0}, chunks = {256, 256, 256};
  double pwe = 0.0, psnr = 0.0, bpp = 0.0;
  int nquality = 0;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    auto need = [&](int n) {
      if (i + n >= argc) {
        std::cout << a << ": " << n << " value(s) required" << std::endl;
        std::exit(106);
      }
    };
    auto num = [&](const char* s) { return size_t(std::strtoull(s, nullptr, 10)); };
    if (a == "-h" || a == "--help") { usage(); return 0; }
    else if (a == "-c") cflag = true;
    else if (a == "-d") dflag = true;
    else if (a == "--omp") { need(1); omp = num(argv[++i]); }
    else if (a == "--ftype") { need(1); ftype = num(argv[++i]); }
    else if (a == "--dims") { need(3); for (int k = 0; k < 3; k++) dims[k] = num(argv[++i]); }
    else if (a == "--chunks") { need(3); for (int k = 0; k < 3; k++) chunks[k] = num(argv[++i]); }
    else if (a == "--bitstream") { need(1); bitstream = argv[++i]; }
    else if (a == "--decomp_f") { need(1); decomp_f32 = argv[++i]; }
    else if (a == "--decomp_d") { need(1); decomp_f64 = argv[++i]; }
    else if (a == "--decomp_lowres_f") { need(1); lowres_f32 = argv[++i]; }
    else if (a == "--decomp_lowres_d") { need(1); lowres_f64 = argv[++i]; }
    else if (a == "--print_stats") print_stats = true;
    else if (a == "--pwe") { need(1); pwe = std::atof(argv[++i]); nquality++; }
    else if (a == "--psnr") { need(1); psnr = std::atof(argv[++i]); nquality++; }
    else if (a == "--bpp") { need(1); bpp = std::atof(argv[++i]); nquality++; }
    else if (!a.empty() && a[0] == '-') {
      std::cout << "The following argument was not expected: " << a << std::endl;
      return 109;
    }
    else
      input_file = a;
  }
  (void)omp;
  // the exclusions CLI11 enforces in the reference
  if (cflag && dflag) { std::cout << "-d excludes -c" << std::endl; return 108; }
  if (nquality > 1) { std::cout << "--pwe, --psnr and --bpp exclude each other" << std::endl; return 108; }
  if (bpp < 0.0 || bpp > 64.0) { std::cout << "--bpp: value not in range 0 to 64" << std::endl; return 105; }
  if ((!bitstream.empty() || print_stats) && !cflag) { std::cout << "--bitstream / --print_stats require -c" << std::endl; return 107; }
  // utilities/sperr3d.cpp:207-262
  if (input_file.empty()) { std::cout << "What's the input file?" << std::endl; return 1; }
  if (!cflag && !dflag) { std::cout << "Is this compressing (-c) or decompressing (-d) ?" << std::endl; return 1; }
  if (cflag && dims == std::array<size_t, 3>{0, 0, 0}) {
    std::cout << "What's the dimensions of this 3D volume (--dims) ?" << std::endl;
    return 1;
  }
  if (cflag && ftype != 32 && ftype != 64) {
    std::cout << "What's the floating-type precision (--ftype) ?" << std::endl;
    return 1;
  }
  if (cflag && pwe == 0.0 && psnr == 0.0 && bpp == 0.0) {
    std::cout << "What's the compression quality (--psnr, --pwe, --bpp) ?" << std::endl;
    return 1;
  }
  if (cflag && (pwe < 0.0 || psnr < 0.0)) {
    std::cout << "Compression quality (--psnr, --pwe) must be positive!" << std::endl;
    return 1;
  }
  const bool multi_res = !lowres_f32.empty() || !lowres_f64.empty();
  if (cflag && multi_res && coarsened_resolutions(dims, {std::min(chunks[0], dims[0]), std::min(chunks[1], dims[1]),
                                                          std::min(chunks[2], dims[2])}).empty()) {
    std::printf(
        " Warning: the combo of volume dimension (%lu, %lu, %lu) and chunk dimension"
        " (%lu, %lu, %lu)\n cannot support multi-resolution decoding. "
        " Try to use chunk dimensions that\n are similar in length and"
        " can divide the volume dimension.\n",
        dims[0], dims[1], dims[2], chunks[0], chunks[1], chunks[2]);
    return 1;
  }
  if (dflag && decomp_f32.empty() && decomp_f64.empty() && !multi_res) {
    std::cout << "SPERR needs an output destination when decoding!" << std::endl;
    return 1;
  }
  if (cflag && bitstream.empty())
    std::cout << "Warning: no output file provided. Consider using --bitstream option." << std::endl;

  bool ok = false;
  auto input = read_whole_file(input_file, ok);
  if (!ok) { std::cout << "Cannot read " << input_file << std::endl; return 1; }
  if (cflag) {
    const size_t total_vals = dims[0] * dims[1] * dims[2];
    if ((ftype == 32 && total_vals * 4 != input.size()) || (ftype == 64 && total_vals * 8 != input.size())) {
      std::cout << "Input file size wrong!" << std::endl;
      return 1;
    }
    sp::SPERR3D_OMP_C encoder;
    encoder.set_dims_and_chunks(dims, chunks);
    encoder.set_num_threads(omp);
    if (pwe != 0.0) encoder.set_tolerance(pwe);
    else if (psnr != 0.0) encoder.set_psnr(psnr);
    else encoder.set_bitrate(bpp);
    const auto rtn = ftype == 32 ? encoder.compress(reinterpret_cast<const float*>(input.data()), total_vals)
                                 : encoder.compress(reinterpret_cast<const double*>(input.data()), total_vals);
    if (rtn != sp::RTNType::Good) { std::cout << "Compression failed!" << std::endl; return 1; }
    const auto stream = encoder.get_encoded_bitstream();
    if (!bitstream.empty() && !write_bytes(bitstream, stream.data(), stream.size())) {
      std::cout << "Writing compressed bitstream failed: " << bitstream << std::endl;
      return 1;
    }
    if (print_stats || !decomp_f64.empty() || !decomp_f32.empty() || multi_res) {
      sp::SPERR3D_OMP_D decoder;
      decoder.use_bitstream(stream.data(), stream.size());
      if (decoder.decompress(stream.data(), multi_res) != sp::RTNType::Good) {
        std::cout << "Decompression failed!" << std::endl;
        return 1;
      }
      const auto outputd = decoder.release_decoded_data();
      if (output_hierarchy(decoder.view_hierarchy(), coarsened_resolutions(decoder.get_dims(), decoder.get_chunk_dims()),
                           lowres_f64, lowres_f32))
        return 1;
      if (output_buffer(outputd, decomp_f64, decomp_f32))
        return 1;
      if (print_stats) {
        const double print_bpp = stream.size() * 8.0 / total_vals;
        double rmse, linfy, print_psnr, mn, mx, sigma;
        if (ftype == 32) {
          const float* inputf = reinterpret_cast<const float*>(input.data());
          std::vector<float> outputf(total_vals);
          std::copy(outputd.cbegin(), outputd.cend(), outputf.begin());
          const auto st = calc_stats(inputf, outputf.data(), total_vals);
          rmse = st[0]; linfy = st[1]; print_psnr = st[2]; mn = st[3]; mx = st[4];
          sigma = std::sqrt(calc_mean_var(inputf, total_vals)[1]);
        }
        else {
          const double* inputd = reinterpret_cast<const double*>(input.data());
          const auto st = calc_stats(inputd, outputd.data(), total_vals);
          rmse = st[0]; linfy = st[1]; print_psnr = st[2]; mn = st[3]; mx = st[4];
          sigma = std::sqrt(calc_mean_var(inputd, total_vals)[1]);
        }
        std::printf("Input range = (%.2e, %.2e), L-Infty = %.2e\n", mn, mx, linfy);
        std::printf("Bitrate = %.2f, PSNR = %.2fdB, Accuracy Gain = %.2f\n", print_bpp, print_psnr,
                    std::log2(sigma / rmse) - print_bpp);
      }
    }
  }
  else {
    sp::SPERR3D_OMP_D decoder;
    decoder.use_bitstream(input.data(), input.size());
    if (decoder.decompress(input.data(), multi_res) != sp::RTNType::Good) {
      std::cout << "Decompression failed!" << std::endl;
      return 1;
    }
    const auto outputd = decoder.release_decoded_data();
    if (output_hierarchy(decoder.view_hierarchy(), coarsened_resolutions(decoder.get_dims(), decoder.get_chunk_dims()),
                         lowres_f64, lowres_f32))
      return 1;
    if (output_buffer(outputd, decomp_f64, decomp_f32))
      return 1;
  }
  return 0;
}
