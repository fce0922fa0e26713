// Reference-compatible C API (include/sperr_b200.h section 1) and its device-pointer variants.
// Host side of the container format: header build / parse are byte-exact restatements of
//   SPERR3D_OMP_C::m_generate_header          /root/reference/src/SPERR3D_OMP_C.cpp:163-234
//   SPERR3D_Stream_Tools::get_stream_header   /root/reference/src/SPERR3D_Stream_Tools.cpp:46-105
//   C_API::sperr_comp_3d / sperr_decomp_3d    /root/reference/src/SPERR_C_API.cpp:156-258
#include "../../include/sperr_b200.h"

#include <sys/mman.h>

#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>

#include "hostpipe.h"
#include "pipeline.h"

using namespace sperr_b200;

namespace {

class SerialWorker;
// Per-device state of the entry points (the device is the one the calling host thread has selected):
// pipelines, grow-only device staging of the host-pointer API (input volume, container, output
// volume), the copy stream and its helper thread. One job at a time per device (shared_api_mutex).
struct DevState {
  Compressor* comp = nullptr;
  Decompressor* decomp = nullptr;
  rt::DBuf in, stream, vol, cstream;
#ifndef SPERR_EMUL
  cudaStream_t copy_stream = nullptr;
#endif
  SerialWorker* d2h_worker = nullptr;
};
DevState& dev_state()
{
  static DevState s[rt::kMaxDevices];
  return s[rt::cur_dev()];
}
#define g_mutex (shared_api_mutex())
#define g_comp (dev_state().comp)
#define g_decomp (dev_state().decomp)
#define g_in (dev_state().in)
#define g_stream (dev_state().stream)
#define g_vol (dev_state().vol)
#define g_cstream (dev_state().cstream)
#define g_copy_stream (dev_state().copy_stream)
#define g_d2h_worker (dev_state().d2h_worker)

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

// One helper thread that runs posted jobs in order (device -> host copies of finished batches while
// the next batch is being decoded).
class SerialWorker {
 public:
  ~SerialWorker()
  {
    {
      std::lock_guard<std::mutex> l(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    if (th_.joinable())
      th_.join();
  }
  void post(std::function<void()> fn)
  {
    {
      std::lock_guard<std::mutex> l(mu_);
      if (!th_.joinable())
        th_ = std::thread([this] { run(); });
      q_.push_back(std::move(fn));
      busy_++;
    }
    cv_.notify_all();
  }
  // waits for everything posted so far; rethrows the first failure
  void drain()
  {
    std::unique_lock<std::mutex> l(mu_);
    done_.wait(l, [this] { return busy_ == 0; });
    if (err_) {
      auto e = err_;
      err_ = nullptr;
      std::rethrow_exception(e);
    }
  }

 private:
  void run()
  {
    std::unique_lock<std::mutex> l(mu_);
    for (;;) {
      cv_.wait(l, [this] { return stop_ || !q_.empty(); });
      if (q_.empty())
        return;
      auto fn = std::move(q_.front());
      q_.pop_front();
      l.unlock();
      try {
        fn();
      }
      catch (...) {
        l.lock();
        if (!err_)
          err_ = std::current_exception();
        l.unlock();
      }
      l.lock();
      if (--busy_ == 0)
        done_.notify_all();
    }
  }
  std::thread th_;
  std::mutex mu_;
  std::condition_variable cv_, done_;
  std::deque<std::function<void()>> q_;
  size_t busy_ = 0;
  bool stop_ = false;
  std::exception_ptr err_;
};

// wall-clock phases of the host-pointer entry points, printed when SPERR_B200_TIMING is set
struct PhaseTimer {
  const bool on = std::getenv("SPERR_B200_TIMING") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  std::string line;
  void mark(const char* what)
  {
    if (!on)
      return;
    const auto t1 = std::chrono::steady_clock::now();
    char buf[64];
    std::snprintf(buf, sizeof(buf), " %s %.1f ms", what, std::chrono::duration<double, std::milli>(t1 - t0).count());
    line += buf;
    t0 = t1;
  }
  void done(const char* fn)
  {
    if (on)
      std::fprintf(stderr, "sperr_b200 timing %s:%s\n", fn, line.c_str());
  }
};

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
    g_comp = &shared_compressor();
  SrcVol sv{d_src, is_float, dimx, dimy, dimz};
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
  try {
    HostPipe::get().d2h(o + pos, d_out.p, total - hlen, st);
  }
  catch (...) {
    std::free(o);   // (guarded() turns the exception into -1: the buffer must not outlive it)
    throw;
  }
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
  // untrusted dimensions: the chunk count must fit the bytes that are there before any vector of
  // that size is built
  size_t nseg[3], nchunks = 0;
  if (!chunk_grid(ci.vol, ci.cd, nseg, &nchunks) || nchunks > (len - pos) / 4)
    return false;
  ci.chunks = chunk_volume(ci.vol, ci.cd);
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


void decomp_3d_device(const uint8_t* h_stream, const uint8_t* d_stream, const ContainerInfo& ci,
                      int output_float, void* d_dst, cudaStream_t st)
{
  if (!g_decomp)
    g_decomp = &shared_decompressor();
  SrcVol dv{d_dst, output_float, ci.vol[0], ci.vol[1]};
  g_decomp->decompress(h_stream, d_stream, ci.chunks, ci.cs, dv, st);
}

// Shared tail of the 2D entry points: every slice is one chunk of extent dimx x dimy x 1 of a
// "volume" of nslices planes, coded by SPECK2D_FLT's pipeline (dwt2d + SPECK2D_INT).
std::vector<Chunk> slice_chunks(size_t dimx, size_t dimy, size_t nslices)
{
  if (dimx == 0 || dimy == 0 || nslices == 0 || dimx > 65535 || dimy > 65535 ||
      dimx * dimy >= (1ull << 31) || nslices > 0xFFFFFFFFull)
    throw std::runtime_error("unsupported slice extents");
  std::vector<Chunk> chunks(nslices);
  for (size_t s = 0; s < nslices; s++)
    chunks[s] = Chunk{0, uint32_t(dimx), 0, uint32_t(dimy), uint32_t(s), 1};
  return chunks;
}

int comp_2d_device(const void* d_src, int is_float, size_t dimx, size_t dimy, size_t nslices, int mode,
                   double quality, int inc_header, void** dst, size_t* lens_out, cudaStream_t st)
{
  const auto chunks = slice_chunks(dimx, dimy, nslices);
  if (!g_comp)
    g_comp = &shared_compressor();
  g_comp->max_batch = 0;
  g_comp->before_batch = nullptr;
  SrcVol sv{d_src, is_float, dimx, dimy};
  rt::DBuf& d_out = g_cstream;
  d_out.reserve(size_t(1) << 20);
  std::vector<size_t> lens;
  g_comp->compress(sv, chunks, mode, quality, true, d_out, lens, st);
  const size_t hl = inc_header ? 10 : 0;
  size_t payload = 0;
  for (size_t l : lens)
    payload += l;
  const size_t total = payload + hl * nslices;
  uint8_t* o = static_cast<uint8_t*>(std::malloc(std::max<size_t>(total, 1)));
  if (!o)
    return -1;
  // one device -> host copy into the tail, then every stream moves forward behind its header
  uint8_t* tail = o + hl * nslices;
  try {
    HostPipe::get().d2h(tail, d_out.p, payload, st);
    HostPipe::get().wait_idle();
  }
  catch (...) {
    std::free(o);
    throw;
  }
  size_t src_off = 0, dst_off = 0;
  for (size_t s = 0; s < nslices; s++) {
    if (inc_header) {
      uint8_t* h = o + dst_off;
      h[0] = 0;                              // SPERR_VERSION_MAJOR
      h[1] = uint8_t(is_float ? 0x20 : 0);   // not a portion, 2D, float flag
      const uint32_t d2[2] = {uint32_t(dimx), uint32_t(dimy)};
      std::memmove(o + dst_off + 10, tail + src_off, lens[s]);
      std::memcpy(h + 2, d2, 8);
    }
    lens_out[s] = lens[s] + hl;
    src_off += lens[s];
    dst_off += lens[s] + hl;
  }
  *dst = o;
  return 0;
}

void decomp_2d_device(const uint8_t* h_streams, const size_t* lens, size_t nslices, int output_float,
                      size_t dimx, size_t dimy, void* d_dst, cudaStream_t st)
{
  if (!h_streams || !lens)
    throw std::runtime_error("no stream");
  const auto chunks = slice_chunks(dimx, dimy, nslices);
  std::vector<ChunkStream> cs(nslices);
  size_t off = 0;
  for (size_t s = 0; s < nslices; s++) {
    cs[s].off = off;
    cs[s].len = lens[s];
    off += lens[s];
  }
  g_stream.reserve(off + 16);
  HostPipe::get().h2d(g_stream.p, h_streams, off, st);
  if (!g_decomp)
    g_decomp = &shared_decompressor();
  g_decomp->max_batch = 0;
  g_decomp->after_batch = nullptr;
  SrcVol dv{d_dst, output_float, dimx, dimy};
  g_decomp->decompress(h_streams, g_stream.as<uint8_t>(), chunks, cs, dv, st, true);
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
  {   // several GPUs (SPERR_B200_DEVICES): the chunks of this call are spread over them
    const std::vector<int> devs = multi_devices();
    if (!devs.empty() && dimx && dimy && dimz) {
      const size_t vol[3] = {dimx, dimy, dimz}, ck[3] = {chunk_x, chunk_y, chunk_z};
      const int rc = comp_3d_multi(src, is_float, vol, ck, mode, quality, devs, dst, dst_len);
      if (rc != -2)
        return rc;
    }
  }
  if (dimx && dimy && dimz) {   // a volume that does not fit in HBM beside the work buffers: slab groups
    const size_t vol[3] = {dimx, dimy, dimz}, ck[3] = {chunk_x, chunk_y, chunk_z};
    const int rc = comp_3d_streamed(src, is_float, vol, ck, mode, quality, dst, dst_len);
    if (rc != -2)
      return rc;
  }
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t bytes = dimx * dimy * dimz * (is_float ? 4 : 8);
    PhaseTimer pt;
    g_in.reserve(bytes);
    if (!g_comp)
      g_comp = &shared_compressor();
    g_comp->max_batch = 0;
    g_comp->before_batch = nullptr;
#ifndef SPERR_EMUL
    // Pinned source and several z-slabs of chunks: upload slab groups on a copy stream and let the
    // coder start on the first group while the rest is still in flight.
    std::vector<cudaEvent_t> evs;
    const size_t esz = is_float ? 4 : 8;
    if (rt::is_pinned_host(src) && dimx && dimy && dimz) {
      const size_t vol[3] = {dimx, dimy, dimz};
      size_t cd[3] = {chunk_x, chunk_y, chunk_z};
      for (int i = 0; i < 3; i++)
        cd[i] = std::min(std::max<size_t>(1, cd[i]), vol[i]);
      const auto chunks = chunk_volume(vol, cd);
      size_t per_slab = 0;
      while (per_slab < chunks.size() && chunks[per_slab].z0 == chunks[0].z0)
        per_slab++;
      const size_t nslabs = chunks.size() / std::max<size_t>(per_slab, 1);
      // worth it only when every group still fills the GPU: the coder's plane loop and the
      // per-chunk kernels are latency-bound, so small groups cost more than the overlap saves
      // measured at 1024^3 (64 chunks, pinned source): one batch 140 ms, batches of 32 chunks 121 ms,
      // of 16 chunks 118 ms -- the coder of a batch runs while the next batch is still uploading
      const size_t min_group = std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")
                                   ? size_t(std::atoi(std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")))
                                   : 16;
      if (nslabs >= 2 && per_slab * nslabs == chunks.size() && bytes >= (size_t(256) << 20) &&
          chunks.size() >= 2 * min_group) {
        const size_t groups = std::min<size_t>(nslabs, std::min<size_t>(4, chunks.size() / min_group));
        const size_t slabs_per = (nslabs + groups - 1) / groups;
        g_comp->max_batch = slabs_per * per_slab;
        if (!g_copy_stream)
          RT_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        const size_t plane = dimx * dimy * esz;
        for (size_t s0 = 0; s0 < nslabs; s0 += slabs_per) {
          const size_t s1 = std::min(nslabs, s0 + slabs_per);
          const size_t z0 = chunks[s0 * per_slab].z0;
          const size_t z1 = s1 == nslabs ? dimz : chunks[s1 * per_slab].z0;
          rt::h2d(static_cast<char*>(g_in.p) + z0 * plane, static_cast<const char*>(src) + z0 * plane,
                  (z1 - z0) * plane, g_copy_stream);
          cudaEvent_t e;
          RT_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          RT_CHECK(cudaEventRecord(e, g_copy_stream));
          evs.push_back(e);
        }
        const size_t mb = g_comp->max_batch;
        g_comp->before_batch = [&evs, mb, st](size_t first, size_t count) {
          for (size_t g = first / mb; g <= (first + count - 1) / mb && g < evs.size(); g++)
            RT_CHECK(cudaStreamWaitEvent(st, evs[g], 0));
        };
      }
      else if (nslabs >= 2 && per_slab * nslabs == chunks.size() && bytes >= (size_t(256) << 20) &&
               !std::getenv("SPERR_B200_NO_GROUP_H2D")) {
        // One batch: upload slab groups on the copy stream; statistics and forward transform of a
        // group start when its slabs have arrived (Compressor::before_group), the coder runs once
        // over all chunks afterwards.
        const size_t groups = std::min<size_t>(nslabs, 4);
        const size_t slabs_per = (nslabs + groups - 1) / groups;
        const size_t gc = slabs_per * per_slab;
        if (!g_copy_stream)
          RT_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        const size_t plane = dimx * dimy * esz;
        for (size_t s0 = 0; s0 < nslabs; s0 += slabs_per) {
          const size_t s1 = std::min(nslabs, s0 + slabs_per);
          const size_t z0 = chunks[s0 * per_slab].z0;
          const size_t z1 = s1 == nslabs ? dimz : chunks[s1 * per_slab].z0;
          rt::h2d(static_cast<char*>(g_in.p) + z0 * plane, static_cast<const char*>(src) + z0 * plane,
                  (z1 - z0) * plane, g_copy_stream);
          cudaEvent_t e;
          RT_CHECK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
          RT_CHECK(cudaEventRecord(e, g_copy_stream));
          evs.push_back(e);
        }
        g_comp->group_chunks = gc;
        g_comp->before_group = [&evs, gc, st](size_t first, size_t count) {
          for (size_t g = first / gc; g <= (first + count - 1) / gc && g < evs.size(); g++)
            RT_CHECK(cudaStreamWaitEvent(st, evs[g], 0));
        };
      }
    }
    // Pageable source (what an unchanged C caller passes) and several z-slabs of chunks: a helper
    // thread takes the slab groups through the pinned ring (HostPipe::h2d) one after the other, the
    // coder starts on a group as soon as it has arrived. Measured before this existed: upload 88 ms,
    // THEN coder 58 ms, whatever the group size (the overlap above needs a pinned source).
    struct Uploader {
      std::thread th;
      std::mutex mu;
      std::condition_variable cv;
      size_t done = 0;
      std::exception_ptr err;
      ~Uploader()
      {
        if (th.joinable())
          th.join();
      }
    } up;
    if (evs.empty() && !rt::is_pinned_host(src) && dimx && dimy && dimz && !std::getenv("SPERR_B200_NO_PAGEABLE_OVERLAP")) {
      const size_t vol[3] = {dimx, dimy, dimz};
      size_t cd[3] = {chunk_x, chunk_y, chunk_z};
      for (int i = 0; i < 3; i++)
        cd[i] = std::min(std::max<size_t>(1, cd[i]), vol[i]);
      const auto chunks = chunk_volume(vol, cd);
      size_t per_slab = 0;
      while (per_slab < chunks.size() && chunks[per_slab].z0 == chunks[0].z0)
        per_slab++;
      const size_t nslabs = chunks.size() / std::max<size_t>(per_slab, 1);
      const size_t min_group = std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")
                                   ? size_t(std::atoi(std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")))
                                   : 16;
      if (nslabs >= 2 && per_slab * nslabs == chunks.size() && bytes >= (size_t(256) << 20) &&
          chunks.size() >= 2 * min_group) {
        const size_t groups = std::min<size_t>(nslabs, std::min<size_t>(4, chunks.size() / min_group));
        const size_t slabs_per = (nslabs + groups - 1) / groups;
        const size_t mb = slabs_per * per_slab;
        g_comp->max_batch = mb;
        if (!g_copy_stream)
          RT_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        int dev = 0;
        RT_CHECK(cudaGetDevice(&dev));
        const size_t plane = dimx * dimy * esz;
        std::vector<std::pair<size_t, size_t>> parts;   // byte range of every group
        for (size_t s0 = 0; s0 < nslabs; s0 += slabs_per) {
          const size_t s1 = std::min(nslabs, s0 + slabs_per);
          const size_t z0 = chunks[s0 * per_slab].z0;
          const size_t z1 = s1 == nslabs ? dimz : chunks[s1 * per_slab].z0;
          parts.emplace_back(z0 * plane, (z1 - z0) * plane);
        }
        void* const d_in = g_in.p;
        cudaStream_t cs = g_copy_stream;
        up.th = std::thread([&up, parts, d_in, src, dev, cs] {
          try {
            RT_CHECK(cudaSetDevice(dev));
            for (size_t g = 0; g < parts.size(); g++) {
              HostPipe::get().h2d(static_cast<char*>(d_in) + parts[g].first,
                                  static_cast<const char*>(src) + parts[g].first, parts[g].second, cs);
              {
                std::lock_guard<std::mutex> l(up.mu);
                up.done = g + 1;
              }
              up.cv.notify_all();
            }
          }
          catch (...) {
            {
              std::lock_guard<std::mutex> l(up.mu);
              up.err = std::current_exception();
            }
            up.cv.notify_all();
          }
        });
        g_comp->before_batch = [&up, mb](size_t first, size_t count) {
          const size_t need = (first + count - 1) / mb + 1;
          std::unique_lock<std::mutex> l(up.mu);
          up.cv.wait(l, [&] { return up.done >= need || up.err; });
          if (up.err)
            std::rethrow_exception(up.err);
        };
      }
    }
    struct EvGuard {
      std::vector<cudaEvent_t>& v;
      ~EvGuard()
      {
        if (!v.empty() && g_copy_stream)   // (also on the exception path: uploads from the caller's buffer
          cudaStreamSynchronize(g_copy_stream);   // must not be in flight when the call returns)
        for (auto e : v)
          cudaEventDestroy(e);
        if (g_comp) {
          g_comp->max_batch = 0;
          g_comp->before_batch = nullptr;
          g_comp->group_chunks = 0;
          g_comp->before_group = nullptr;
        }
      }
    } ev_guard{evs};
    if (evs.empty() && !up.th.joinable())
#endif
      HostPipe::get().h2d(g_in.p, src, bytes, st);
    pt.mark("h2d");
    const int rc = comp_3d_device(g_in.p, is_float, dimx, dimy, dimz, chunk_x, chunk_y, chunk_z, mode,
                          quality, dst, dst_len, st);
    pt.mark("compress+d2h");
    pt.done("sperr_comp_3d");
    return rc;
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
  {
    const std::vector<int> devs = multi_devices();
    if (!devs.empty() && src && src_len >= 14) {
      const int rc = decomp_3d_multi(src, src_len, output_float, devs, dimx, dimy, dimz, dst);
      if (rc != -2)
        return rc;
    }
  }
  if (src && src_len >= 14) {   // a volume that does not fit in HBM beside the work buffers: slab groups
    const int rc = decomp_3d_streamed(src, src_len, output_float, dimx, dimy, dimz, dst);
    if (rc != -2)
      return rc;
  }
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
    PhaseTimer pt;
    HostPipe::get().prefault_begin(o, total * esz);
    if (!g_decomp)
      g_decomp = &shared_decompressor();
    g_decomp->max_batch = 0;
    g_decomp->after_batch = nullptr;
    bool batched = false;
    try {
      g_stream.reserve(src_len);
      HostPipe::get().h2d(g_stream.p, src, src_len, st);
      pt.mark("h2d");
      g_vol.reserve(total * esz);
#ifndef SPERR_EMUL
      // Several z-slabs of chunks: decode them group by group and copy every finished group to the
      // host (helper thread + copy stream) while the next one is being decoded.
      size_t per_slab = 0;
      while (per_slab < ci.chunks.size() && ci.chunks[per_slab].z0 == ci.chunks[0].z0)
        per_slab++;
      const size_t nslabs = ci.chunks.size() / std::max<size_t>(per_slab, 1);
      const size_t min_group = std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")
                                   ? size_t(std::atoi(std::getenv("SPERR_B200_OVERLAP_MIN_CHUNKS")))
                                   : 256;   // one CTA per chunk stream: smaller groups leave SMs idle
      if (nslabs >= 2 && per_slab * nslabs == ci.chunks.size() && total * esz >= (size_t(256) << 20) &&
          ci.chunks.size() >= 2 * min_group) {
        batched = true;
        const size_t groups = std::min<size_t>(nslabs, std::min<size_t>(4, ci.chunks.size() / min_group));
        const size_t slabs_per = (nslabs + groups - 1) / groups;
        g_decomp->max_batch = slabs_per * per_slab;
        if (!g_copy_stream)
          RT_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        if (!g_d2h_worker)
          g_d2h_worker = new SerialWorker();
        int dev = 0;
        RT_CHECK(cudaGetDevice(&dev));
        const size_t plane = ci.vol[0] * ci.vol[1] * esz;
        const auto& chunks = ci.chunks;
        const size_t dimz = ci.vol[2];
        g_decomp->after_batch = [=, &chunks](size_t first, size_t count) {
          // the batch is complete in g_vol (run_batch synchronises): its z range is final when the
          // batch ends on a slab boundary
          const size_t z0 = chunks[first].z0;
          const size_t end = first + count;
          const size_t z1 = end == chunks.size() ? dimz : chunks[end].z0;
          char* hdst = static_cast<char*>(o) + z0 * plane;
          const char* dsrc = static_cast<const char*>(g_vol.p) + z0 * plane;
          const size_t nbytes = (z1 - z0) * plane;
          g_d2h_worker->post([=] {
            RT_CHECK(cudaSetDevice(dev));
            HostPipe::get().d2h(hdst, dsrc, nbytes, g_copy_stream);
          });
        };
      }
#endif
#ifndef SPERR_EMUL
      // One batch, several z-slabs of chunks: the last stage of the batch runs slab group by slab
      // group and every finished group starts its way to the host while the next is transformed.
      if (!batched && nslabs >= 2 && per_slab * nslabs == ci.chunks.size() &&
          total * esz >= (size_t(256) << 20) && !std::getenv("SPERR_B200_NO_GROUP_D2H")) {
        const size_t groups = std::min<size_t>(nslabs, 4);
        const size_t slabs_per = (nslabs + groups - 1) / groups;
        g_decomp->group_chunks = slabs_per * per_slab;
        if (!g_copy_stream)
          RT_CHECK(cudaStreamCreateWithFlags(&g_copy_stream, cudaStreamNonBlocking));
        if (!g_d2h_worker)
          g_d2h_worker = new SerialWorker();
        int dev = 0;
        RT_CHECK(cudaGetDevice(&dev));
        const size_t plane = ci.vol[0] * ci.vol[1] * esz;
        const auto& chunks = ci.chunks;
        const size_t dimz = ci.vol[2];
        g_decomp->after_group = [=, &chunks](size_t first, size_t count, cudaEvent_t ev) {
          const size_t z0 = chunks[first].z0;
          const size_t end = first + count;
          const size_t z1 = end == chunks.size() ? dimz : chunks[end].z0;
          char* hdst = static_cast<char*>(o) + z0 * plane;
          const char* dsrc = static_cast<const char*>(g_vol.p) + z0 * plane;
          const size_t nbytes = (z1 - z0) * plane;
          g_d2h_worker->post([=] {
            RT_CHECK(cudaSetDevice(dev));
            RT_CHECK(cudaStreamWaitEvent(g_copy_stream, ev, 0));
            HostPipe::get().d2h(hdst, dsrc, nbytes, g_copy_stream);
            cudaEventDestroy(ev);
          });
        };
      }
#endif
      decomp_3d_device(static_cast<const uint8_t*>(src), g_stream.as<uint8_t>(), ci, output_float,
                       g_vol.p, st);
      pt.mark("decode");
#ifndef SPERR_EMUL
      if (!batched && g_decomp->groups_posted)
        batched = true;   // the result is already on its way: wait for the copies below
      g_decomp->group_chunks = 0;
      g_decomp->after_group = nullptr;
#endif
      if (batched) {
#ifndef SPERR_EMUL
        g_d2h_worker->drain();
#endif
        pt.mark("d2h tail");
      }
      else {
        HostPipe::get().d2h(o, g_vol.p, total * esz, st);
        pt.mark("d2h");
      }
      HostPipe::get().wait_idle();
      g_decomp->max_batch = 0;
      g_decomp->after_batch = nullptr;
      pt.done("sperr_decomp_3d");
    }
    catch (...) {
#ifndef SPERR_EMUL
      if (batched && g_d2h_worker) {
        try {
          g_d2h_worker->drain();
        }
        catch (...) {
        }
      }
#endif
      HostPipe::get().wait_idle();
      g_decomp->max_batch = 0;
      g_decomp->after_batch = nullptr;
      g_decomp->group_chunks = 0;
      g_decomp->after_group = nullptr;
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

// ---- 2D slices (C_API::sperr_comp_2d / sperr_decomp_2d, /root/reference/src/SPERR_C_API.cpp:11-134) ----

int sperr_b200_comp_2d_batch_dev(const void* d_src, int is_float, size_t dimx, size_t dimy,
                                 size_t nslices, int mode, double quality, int out_inc_header,
                                 void** dst, size_t* lens)
{
  if (*dst != nullptr)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] { return comp_2d_device(d_src, is_float, dimx, dimy, nslices, mode, quality,
                                             out_inc_header, dst, lens, 0); });
}

int sperr_b200_comp_2d_batch(const void* src, int is_float, size_t dimx, size_t dimy, size_t nslices,
                             int mode, double quality, int out_inc_header, void** dst, size_t* lens)
{
  if (*dst != nullptr)
    return 1;
  if (quality <= 0.0)
    return 2;
  if (mode < 1 || mode > 3)
    return 2;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t bytes = dimx * dimy * nslices * (is_float ? 4 : 8);
    g_in.reserve(bytes);
    HostPipe::get().h2d(g_in.p, src, bytes, st);
    return comp_2d_device(g_in.p, is_float, dimx, dimy, nslices, mode, quality, out_inc_header, dst,
                          lens, st);
  });
}

int sperr_comp_2d(const void* src, int is_float, size_t dimx, size_t dimy, int mode, double quality,
                  int out_inc_header, void** dst, size_t* dst_len)
{
  size_t len = 0;
  const int rc = sperr_b200_comp_2d_batch(src, is_float, dimx, dimy, 1, mode, quality, out_inc_header,
                                          dst, &len);
  if (rc == 0)
    *dst_len = len;
  return rc;
}

int sperr_b200_decomp_2d_batch_dev(const void* src, const size_t* lens, size_t nslices,
                                   int output_float, size_t dimx, size_t dimy, void* d_dst)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    decomp_2d_device(static_cast<const uint8_t*>(src), lens, nslices, output_float, dimx, dimy, d_dst, st);
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_decomp_2d_batch(const void* src, const size_t* lens, size_t nslices, int output_float,
                               size_t dimx, size_t dimy, void** dst)
{
  if (*dst != nullptr)
    return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    const size_t total = dimx * dimy * nslices, esz = output_float ? 4 : 8;
    if (total == 0)
      return -1;
    g_vol.reserve(total * esz);
    decomp_2d_device(static_cast<const uint8_t*>(src), lens, nslices, output_float, dimx, dimy, g_vol.p, st);
    void* o = std::malloc(total * esz);
    if (!o)
      return -1;
    try {
      HostPipe::get().d2h(o, g_vol.p, total * esz, st);
      HostPipe::get().wait_idle();
    }
    catch (...) {
      std::free(o);
      throw;
    }
    *dst = o;
    return 0;
  });
}

int sperr_decomp_2d(const void* src, size_t src_len, int output_float, size_t dimx, size_t dimy,
                    void** dst)
{
  return sperr_b200_decomp_2d_batch(src, &src_len, 1, output_float, dimx, dimy, dst);
}

// Multi-resolution decoding (SPERR3D_OMP_D::decompress(p, true), /root/reference/src/SPERR3D_OMP_D.cpp:
// 51-135; sperr::coarsened_resolutions, src/sperr_helper.cpp:70-123). The reference's C API has no
// entry for it; this is what the class mirror and tools/sperr3d use.
int sperr_b200_decomp_3d_multires(const void* src, size_t src_len, int output_float, size_t* dimx,
                                  size_t* dimy, size_t* dimz, void** dst, size_t* nlevels,
                                  size_t* level_dims, void** level_data)
{
  if (*dst != nullptr)
    return 1;
  std::lock_guard<std::mutex> lock(g_mutex);
  return guarded([&] {
    cudaStream_t st = 0;
    ContainerInfo ci;
    if (!parse_container(static_cast<const uint8_t*>(src), src_len, ci))
      return -1;
    const size_t esz = output_float ? 4 : 8;
    // levels exist when the volume is a whole number of dyadic chunks
    MultiRes mr;
    mr.is_float = output_float;
    bool divisible = true;
    for (int i = 0; i < 3; i++)
      divisible &= ci.vol[i] % ci.cd[i] == 0;
    const int L = (divisible && ci.cd[2] > 1) ? can_use_dyadic(ci.cd[0], ci.cd[1], ci.cd[2]) : -1;
    std::vector<rt::DBuf> bufs;
    if (L >= 1 && L <= 8 && !BatchBuffers::no_fused()) {
      bufs.resize(size_t(L));
      for (int lev = L; lev >= 1; lev--) {
        std::array<size_t, 3> d;
        for (int i = 0; i < 3; i++)
          d[i] = calc_approx_detail_len(ci.cd[i], size_t(lev))[0] * (ci.vol[i] / ci.cd[i]);
        mr.dims.push_back(d);
        bufs[size_t(L - lev)].alloc(d[0] * d[1] * d[2] * esz);
        mr.d_level.push_back(bufs[size_t(L - lev)].p);
      }
    }
    const size_t total = ci.vol[0] * ci.vol[1] * ci.vol[2];
    g_stream.reserve(src_len);
    HostPipe::get().h2d(g_stream.p, src, src_len, st);
    g_vol.reserve(total * esz);
    if (!g_decomp)
      g_decomp = &shared_decompressor();
    g_decomp->max_batch = 0;
    g_decomp->after_batch = nullptr;
    g_decomp->multires = mr.d_level.empty() ? nullptr : &mr;
    try {
      decomp_3d_device(static_cast<const uint8_t*>(src), g_stream.as<uint8_t>(), ci, output_float, g_vol.p, st);
    }
    catch (...) {
      g_decomp->multires = nullptr;
      throw;
    }
    g_decomp->multires = nullptr;
    std::vector<void*> outs;
    auto fail = [&] {
      for (void* p : outs)
        std::free(p);
      return -1;
    };
    void* o = std::malloc(std::max<size_t>(total * esz, 1));
    if (!o)
      return fail();
    outs.push_back(o);
    try {
      HostPipe::get().d2h(o, g_vol.p, total * esz, st);
      for (size_t h = 0; h < mr.d_level.size(); h++) {
        const size_t nb = mr.dims[h][0] * mr.dims[h][1] * mr.dims[h][2] * esz;
        void* p = std::malloc(std::max<size_t>(nb, 1));
        if (!p)
          return fail();
        outs.push_back(p);
        HostPipe::get().d2h(p, mr.d_level[h], nb, st);
      }
      HostPipe::get().wait_idle();
    }
    catch (...) {
      fail();   // a failed copy must not leak the buffers handed out so far
      throw;
    }
    *dst = o;
    *dimx = ci.vol[0]; *dimy = ci.vol[1]; *dimz = ci.vol[2];
    *nlevels = mr.d_level.size();
    for (size_t h = 0; h < mr.d_level.size(); h++) {
      for (int i = 0; i < 3; i++)
        level_dims[3 * h + i] = mr.dims[h][i];
      level_data[h] = outs[h + 1];
    }
    return 0;
  });
}

// C_API::sperr_trunc_3d (/root/reference/src/SPERR_C_API.cpp:260-281) over
// SPERR3D_Stream_Tools::progressive_truncate / m_progressive_helper
// (src/SPERR3D_Stream_Tools.cpp:131-226): keep the first pct % of every chunk's stream (at least 64
// bytes, or the whole chunk when it is shorter), flag the container as a portion. Host bytes only.
int sperr_trunc_3d(const void* src, size_t src_len, unsigned pct, void** dst, size_t* dst_len)
{
  if (*dst != nullptr)
    return 1;
  const uint8_t* p = static_cast<const uint8_t*>(src);
  if (!p || src_len < 20)   // the reference reads 20 header bytes unconditionally
    return -1;
  const bool multi = (p[1] & 0x10) != 0;
  uint32_t v3[3];
  std::memcpy(v3, p + 2, 12);
  size_t vol[3] = {v3[0], v3[1], v3[2]}, cd[3] = {v3[0], v3[1], v3[2]};
  size_t pos = 14;
  if (multi) {
    uint16_t c3[3];
    std::memcpy(c3, p + 14, 6);
    for (int i = 0; i < 3; i++)
      cd[i] = c3[i];
    pos = 20;
  }
  for (int i = 0; i < 3; i++)
    if (vol[i] == 0 || cd[i] == 0)
      return -1;
  size_t nseg[3], nchunks = 0;
  if (!chunk_grid(vol, cd, nseg, &nchunks) || nchunks > (src_len - pos) / 4)
    return -1;
  const size_t hlen = pos + 4 * nchunks;
  std::vector<size_t> off, len;
  try {
    off.resize(nchunks);
    len.resize(nchunks);
  }
  catch (const std::exception&) {
    return -1;
  }
  size_t at = hlen, far = 0, total = hlen;
  const size_t min_bytes = 64;   // m_progressive_min_chunk_bytes
  for (size_t i = 0; i < nchunks; i++) {
    uint32_t l;
    std::memcpy(&l, p + pos + 4 * i, 4);
    off[i] = at;
    at += l;
    size_t keep = l;
    if (pct != 0 && pct < 100 && keep > min_bytes)
      keep = std::max(min_bytes, size_t(double(pct) / 100.0 * double(keep)));
    len[i] = keep;
    far = std::max(far, off[i] + keep);
    total += keep;
  }
  if (src_len < far)
    return -1;
  uint8_t* o = static_cast<uint8_t*>(std::malloc(total));
  if (!o)
    return -1;
  std::memcpy(o, p, hlen);
  if (pct != 0 && pct < 100) {
    o[0] = 0;       // SPERR_VERSION_MAJOR
    o[1] |= 0x80;   // is_portion
    for (size_t i = 0; i < nchunks; i++) {
      const uint32_t l = uint32_t(len[i]);
      std::memcpy(o + pos + 4 * i, &l, 4);
    }
  }
  size_t w = hlen;
  for (size_t i = 0; i < nchunks; i++) {
    std::memcpy(o + w, p + off[i], len[i]);
    w += len[i];
  }
  *dst = o;
  *dst_len = total;
  return 0;
}

// Arithmetic flavour of the wavelet lifting steps for all later calls (kernels.h: lift_add).
void sperr_b200_set_fma_flavour(int on)
{
  std::lock_guard<std::mutex> lock(g_mutex);
  fma_flavour() = on ? 1 : 0;
}

void sperr_b200_prof_enable(int on)
{
  rt::prof().on = on != 0;
  if (on) {
    rt::prof_collect();   // ranges left over from an earlier session hand their events back
    rt::prof().acc.clear();
    rt::prof().host.clear();
#ifndef SPERR_EMUL
    rt::prof_reserve(4096);   // outside any timed region: the ranges of a timed loop only take from the pool
#endif
  }
}

size_t sperr_b200_prof_dump(char* buf, size_t cap)
{
  rt::prof_collect();
  std::string s = "{";
  bool first = true;
  for (auto& kv : rt::prof().acc) {
    char tmp[256];
    const auto hit = rt::prof().host.find(kv.first);
    const double hsum = hit == rt::prof().host.end() ? 0.0 : hit->second.first;
    const double hmax = hit == rt::prof().host.end() ? 0.0 : hit->second.second;
    std::snprintf(tmp, sizeof(tmp), "%s\"%s\": {\"ms\": %.6f, \"n\": %ld, \"host_ms\": %.3f, \"host_max_ms\": %.3f}",
                  first ? "" : ", ", kv.first.c_str(), kv.second.first, kv.second.second, hsum, hmax);
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
