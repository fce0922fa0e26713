// Reference-compatible C API (include/sperr_b200.h section 1) and its device-pointer variants.
// Host side of the container format: header build / parse are byte-exact restatements of
//   SPERR3D_OMP_C::m_generate_header          /root/reference/src/SPERR3D_OMP_C.cpp:163-234
//   SPERR3D_Stream_Tools::get_stream_header   /root/reference/src/SPERR3D_Stream_Tools.cpp:46-105
//   C_API::sperr_comp_3d / sperr_decomp_3d    /root/reference/src/SPERR_C_API.cpp:156-258
#include "../../include/sperr_b200.h"

#include <sys/mman.h>

#include <mutex>

#include "hostpipe.h"
#include "pipeline.h"

using namespace sperr_b200;

namespace {

std::mutex g_mutex;  // one job at a time per process: the work buffers are shared
Compressor* g_comp = nullptr;
// grow-only device staging of the host-pointer entry points (input volume, container, output volume)
rt::DBuf g_in, g_stream, g_vol, g_cstream;

bool device_ok()
{
#ifndef SPERR_EMUL
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return false;
  }
#endif
  return true;
}

template <typename F>
int guarded(F&& f)
{
  try {
    if (!device_ok())
      throw std::runtime_error("no CUDA device: this library has no CPU path");
    return f();
  }
  catch (const std::exception& e) {
    if (std::getenv("SPERR_B200_VERBOSE"))
      std::fprintf(stderr, "sperr_b200: %s\n", e.what());
    return -1;
  }
}

// Shared tail of sperr_comp_3d: `d_src` is the device-resident volume.
int comp_3d_device(const void* d_src, int is_float, size_t dimx, size_t dimy, size_t dimz,
                   size_t chunk_x, size_t chunk_y, size_t chunk_z, int mode, double quality,
                   void** dst, size_t* dst_len, cudaStream_t st)
{
  const size_t vol[3] = {dimx, dimy, dimz};
  size_t cd[3] = {chunk_x, chunk_y, chunk_z};
  for (int i = 0; i < 3; i++)  // SPERR3D_OMP_C::set_dims_and_chunks, :22-29
    cd[i] = std::min(std::max<size_t>(1, cd[i]), vol[i]);
  if (dimx == 0 || dimy == 0 || dimz == 0)
    return -1;
  if (dimx > 0xFFFFFFFFull || dimy > 0xFFFFFFFFull || dimz > 0xFFFFFFFFull)
    return -1;
  const auto chunks = chunk_volume(vol, cd);
  for (auto& c : chunks)
    if (c.lx > 65535 || c.ly > 65535 || c.lz > 65535 || c.nelem() >= (1ull << 31))
      return -1;  // Set3D coordinates are 16-bit in the reference as well
  if (!g_comp)
    g_comp = new Compressor();
  SrcVol sv{d_src, is_float, dimx, dimy};
  rt::DBuf& d_out = g_cstream;
  d_out.reserve(size_t(1) << 20);
  std::vector<size_t> lens;
  g_comp->compress(sv, chunks, mode, quality, false, d_out, lens, st);

  const size_t nchunks = chunks.size();
  const size_t hlen = (nchunks > 1 ? 20 : 14) + 4 * nchunks;
  size_t total = hlen;
  for (size_t l : lens) {
    if (l > 0xFFFFFFFFull)
      return -1;
    total += l;
  }
  uint8_t* o = static_cast<uint8_t*>(std::malloc(total));
  if (!o)
    return -1;
  o[0] = 0;  // SPERR_VERSION_MAJOR
  o[1] = uint8_t(0x40 | (is_float ? 0x20 : 0) | (nchunks > 1 ? 0x10 : 0));
  const uint32_t v3[3] = {uint32_t(dimx), uint32_t(dimy), uint32_t(dimz)};
  std::memcpy(o + 2, v3, 12);
  size_t pos = 14;
  if (nchunks > 1) {
    const uint16_t c3[3] = {uint16_t(cd[0]), uint16_t(cd[1]), uint16_t(cd[2])};
    std::memcpy(o + pos, c3, 6);
    pos += 6;
  }
  for (size_t l : lens) {
    const uint32_t l32 = uint32_t(l);
    std::memcpy(o + pos, &l32, 4);
    pos += 4;
  }
  HostPipe::get().d2h(o + pos, d_out.p, total - hlen, st);
  *dst = o;
  *dst_len = total;
  return 0;
}

struct ContainerInfo {
  size_t vol[3], cd[3];
  std::vector<Chunk> chunks;
  std::vector<ChunkStream> cs;
};

// SPERR3D_Stream_Tools::get_stream_header (src/SPERR3D_Stream_Tools.cpp:46-105) and the checks of
// SPERR3D_OMP_D::use_bitstream (src/SPERR3D_OMP_D.cpp:23-49): version, 3D flag, total length.
bool parse_container(const uint8_t* p, size_t len, ContainerInfo& ci)
{
  if (!p || len < 14)
    return false;
  if (p[0] != 0)  // SPERR_VERSION_MAJOR
    return false;
  if (!(p[1] & 0x40))
    return false;
  const bool multi = (p[1] & 0x10) != 0;
  uint32_t v3[3];
  std::memcpy(v3, p + 2, 12);
  size_t pos = 14;
  for (int i = 0; i < 3; i++)
    ci.vol[i] = ci.cd[i] = v3[i];
  if (multi) {
    if (len < 20)
      return false;
    uint16_t c3[3];
    std::memcpy(c3, p + 14, 6);
    for (int i = 0; i < 3; i++)
      ci.cd[i] = c3[i];
    pos = 20;
  }
  for (int i = 0; i < 3; i++)
    if (ci.vol[i] == 0 || ci.cd[i] == 0)
      return false;
  ci.chunks = chunk_volume(ci.vol, ci.cd);
  const size_t nchunks = ci.chunks.size();
  if (len < pos + 4 * nchunks)
    return false;
  ci.cs.resize(nchunks);
  size_t off = pos + 4 * nchunks;
  for (size_t i = 0; i < nchunks; i++) {
    uint32_t l;
    std::memcpy(&l, p + pos + 4 * i, 4);
    ci.cs[i].off = off;
    ci.cs[i].len = l;
    off += l;
  }
  if (off != len)
    return false;
  for (auto& c : ci.chunks)
    if (c.nelem() >= (1ull << 31))
      return false;
  return true;
}

Decompressor* g_decomp = nullptr;

void decomp_3d_device(const uint8_t* h_stream, const uint8_t* d_stream, const ContainerInfo& ci,
                      int output_float, void* d_dst, cudaStream_t st)
{
  if (!g_decomp)
    g_decomp = new Decompressor();
  SrcVol dv{d_dst, output_float, ci.vol[0], ci.vol[1]};
  g_decomp->decompress(h_stream, d_stream, ci.chunks, ci.cs, dv, st);
}

}  // namespace

extern "C" {

int sperr_comp_3d(const void* src, int is_float, size_t dimx, size_t dimy, size_t dimz,
                  size_t chunk_x, size_t chunk_y, size_t chunk_z, int mode, double quality,
                  size_t nthreads, void** dst, size_t* dst_len)
{
  (void)nthreads;
  if (*dst != nullptr)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t bytes = dimx * dimy * dimz * (is_float ? 4 : 8);
    g_in.reserve(bytes);
    HostPipe::get().h2d(g_in.p, src, bytes, st);
    return comp_3d_device(g_in.p, is_float, dimx, dimy, dimz, chunk_x, chunk_y, chunk_z, mode,
                          quality, dst, dst_len, st);
  });
}

int sperr_b200_comp_3d_dev(const void* d_src, int is_float, size_t dimx, size_t dimy, size_t dimz,
                           size_t chunk_x, size_t chunk_y, size_t chunk_z, int mode, double quality,
                           void** dst, size_t* dst_len)
{
  if (*dst != nullptr)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    return comp_3d_device(d_src, is_float, dimx, dimy, dimz, chunk_x, chunk_y, chunk_z, mode, quality,
                          dst, dst_len, 0);
  });
}

int sperr_decomp_3d(const void* src, size_t src_len, int output_float, size_t nthreads, size_t* dimx,
                    size_t* dimy, size_t* dimz, void** dst)
{
  (void)nthreads;
  if (*dst != nullptr)
    return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    ContainerInfo ci;
    if (!parse_container(static_cast<const uint8_t*>(src), src_len, ci))
      return -1;
    const size_t total = ci.vol[0] * ci.vol[1] * ci.vol[2];
    const size_t esz = output_float ? 4 : 8;
    // the result buffer first: its pages are faulted in by the copy threads while the GPU decodes
    void* o = std::malloc(total * esz);
    if (!o)
      return -1;
    if (total * esz >= (size_t(64) << 20)) {
      const uintptr_t a = (reinterpret_cast<uintptr_t>(o) + 4095) & ~uintptr_t(4095);
      madvise(reinterpret_cast<void*>(a), (total * esz - (a - reinterpret_cast<uintptr_t>(o))) & ~size_t(4095),
              MADV_HUGEPAGE);
    }
    HostPipe::get().prefault_begin(o, total * esz);
    try {
      g_stream.reserve(src_len);
      HostPipe::get().h2d(g_stream.p, src, src_len, st);
      g_vol.reserve(total * esz);
      decomp_3d_device(static_cast<const uint8_t*>(src), g_stream.as<uint8_t>(), ci, output_float,
                       g_vol.p, st);
      HostPipe::get().wait_idle();
      HostPipe::get().d2h(o, g_vol.p, total * esz, st);
    }
    catch (...) {
      HostPipe::get().wait_idle();
      std::free(o);
      throw;
    }
    *dimx = ci.vol[0];
    *dimy = ci.vol[1];
    *dimz = ci.vol[2];
    *dst = o;
    return 0;
  });
}

int sperr_b200_decomp_3d_dev(const void* h_src, const void* d_src, size_t src_len, int output_float,
                             size_t* dimx, size_t* dimy, size_t* dimz, void* d_dst)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    ContainerInfo ci;
    if (!parse_container(static_cast<const uint8_t*>(h_src), src_len, ci))
      return -1;
    rt::DBuf tmp;
    const uint8_t* ds = static_cast<const uint8_t*>(d_src);
    if (!ds) {
      tmp.alloc(src_len);
      rt::h2d(tmp.p, h_src, src_len, st);
      ds = tmp.as<uint8_t>();
    }
    decomp_3d_device(static_cast<const uint8_t*>(h_src), ds, ci, output_float, d_dst, st);
    rt::sync(st);
    *dimx = ci.vol[0];
    *dimy = ci.vol[1];
    *dimz = ci.vol[2];
    return 0;
  });
}

void sperr_b200_prof_enable(int on)
{
  rt::prof().on = on != 0;
  if (on) {
    rt::prof().acc.clear();
    rt::prof().open.clear();
  }
}

size_t sperr_b200_prof_dump(char* buf, size_t cap)
{
  rt::prof_collect();
  std::string s = "{";
  bool first = true;
  for (auto& kv : rt::prof().acc) {
    char tmp[256];
    std::snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"n\": %ld}", first ? "" : ", ",
                  kv.first.c_str(), kv.second.first, kv.second.second);
    s += tmp;
    first = false;
  }
  s += "}";
  if (buf && cap) {
    const size_t n = std::min(cap - 1, s.size());
    std::memcpy(buf, s.data(), n);
    buf[n] = 0;
  }
  return s.size();
}

unsigned long long sperr_b200_launch_count(void) { return rt::launch_counter().load(); }

void sperr_parse_header(const void* src, size_t* dimx, size_t* dimy, size_t* dimz, int* is_float)
{
  const uint8_t* p = static_cast<const uint8_t*>(src);
  const bool is_3d = (p[1] & 0x40) != 0;
  *is_float = (p[1] & 0x20) ? 1 : 0;
  uint32_t d[3] = {1, 1, 1};
  std::memcpy(d, p + 2, is_3d ? 12 : 8);
  *dimx = d[0];
  *dimy = d[1];
  *dimz = d[2];
}

}  // extern "C"
