// Several GPUs behind the reference's C API: one call of sperr_comp_3d / sperr_decomp_3d spreads its
// chunks over the GPUs named by SPERR_B200_DEVICES ("all", or a list such as "0,1,2,3"), the way the
// reference spreads them over OpenMP threads inside compress() / decompress()
// (/root/reference/src/SPERR3D_OMP_C.cpp:84-161, src/SPERR3D_OMP_D.cpp:51-135; the C API above them:
// src/SPERR_C_API.cpp:181-206, 228-249).
//
// Chunks are independent, so there is no data-path exchange between the GPUs at all: device d gets a
// box of whole chunks (shard_ranges, geom.cpp), one host thread drives it through the chunk-range
// entry points with that device's own pipelines and work buffers, and the only shared objects are
// host-side: the table of chunk lengths the container header needs and the result buffer, into which
// every thread copies its own byte range. Host <-> device traffic of a box goes through a small ring
// of pinned slots per device (the caller's memory is pageable).
#include "../../include/sperr_b200.h"

#include <sys/mman.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <thread>

#include "pipeline.h"

using namespace sperr_b200;

namespace sperr_b200 {

std::vector<int> multi_devices()
{
  std::vector<int> out;
#ifndef SPERR_EMUL
  const char* e = std::getenv("SPERR_B200_DEVICES");
  if (!e || !*e)
    return out;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    return out;
  }
  n = std::min(n, rt::kMaxDevices);
  if (std::string(e) == "all") {
    for (int i = 0; i < n; i++)
      out.push_back(i);
  }
  else {
    const char* p = e;
    while (*p) {
      char* end = nullptr;
      const long v = std::strtol(p, &end, 10);
      if (end == p)
        break;
      if (v >= 0 && v < n && std::find(out.begin(), out.end(), int(v)) == out.end())
        out.push_back(int(v));
      p = *end == ',' ? end + 1 : end;
      if (*end != ',' && *end != 0)
        break;
    }
  }
  if (out.size() < 2)
    out.clear();
#endif
  return out;
}

}  // namespace sperr_b200

#ifndef SPERR_EMUL
namespace {

struct Barrier {   // std::barrier is C++20
  std::mutex mu;
  std::condition_variable cv;
  int n, waiting = 0, gen = 0;
  explicit Barrier(int count) : n(count) {}
  void wait()
  {
    std::unique_lock<std::mutex> l(mu);
    const int g = gen;
    if (++waiting == n) {
      waiting = 0;
      gen++;
      cv.notify_all();
    }
    else
      cv.wait(l, [&] { return gen != g; });
  }
};

// Per-device staging of this file: the device copy of the box and a ring of pinned slots.
constexpr int kRing = 4;
constexpr size_t kSlot = size_t(32) << 20;
struct MultiState {
  rt::DBuf box;
  void* slot[kRing] = {nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev[kRing];
  cudaStream_t st = nullptr;
  void init()
  {
    if (st)
      return;
    RT_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    for (int i = 0; i < kRing; i++) {
      slot[i] = rt::hmalloc_pinned(kSlot);
      RT_CHECK(cudaEventCreateWithFlags(&ev[i], cudaEventDisableTiming));
    }
  }
};
MultiState& multi_state()
{
  static MultiState s[rt::kMaxDevices];
  return s[rt::cur_dev()];
}

// A box of a host volume (x fastest) as a linear byte string: maximal contiguous runs.
struct BoxMap {
  char* base;              // host address of the box's first value
  size_t run, runs_per_plane, run_pitch, plane_pitch, bytes;
  BoxMap(const void* vol, const size_t dims[3], const size_t org[3], const size_t ext[3], size_t esz)
  {
    base = static_cast<char*>(const_cast<void*>(vol)) + ((org[2] * dims[1] + org[1]) * dims[0] + org[0]) * esz;
    plane_pitch = dims[0] * dims[1] * esz;
    if (ext[0] == dims[0]) {   // whole rows: a plane of the box is one run (or the whole box, when whole planes)
      run = ext[0] * ext[1] * esz;
      runs_per_plane = 1;
      run_pitch = 0;
    }
    else {
      run = ext[0] * esz;
      runs_per_plane = ext[1];
      run_pitch = dims[0] * esz;
    }
    bytes = ext[0] * ext[1] * ext[2] * esz;
  }
  char* at(size_t lin, size_t& left_in_run) const
  {
    const size_t r = lin / run, o = lin % run;
    left_in_run = run - o;
    return base + (r / runs_per_plane) * plane_pitch + (r % runs_per_plane) * run_pitch + o;
  }
};

bool pinned(const void* p) { return rt::is_pinned_host(p); }

// Host copies between the caller's (pageable) box and a pinned slot, spread over a few threads: one
// thread moves 5 - 10 GB/s, a PCIe 5 x16 link takes 55. `nthreads` is what the box's cores allow per
// device (all devices stage at the same time).
struct Seg {
  char* dst;
  const char* src;
  size_t n;
};
void par_copy(const std::vector<Seg>& segs, int nthreads)
{
  size_t total = 0;
  for (const Seg& g : segs)
    total += g.n;
  if (nthreads <= 1 || total < (size_t(4) << 20)) {
    for (const Seg& g : segs)
      std::memcpy(g.dst, g.src, g.n);
    return;
  }
  auto share = [&](int k) {   // bytes [lo, hi) of the concatenated segments
    const size_t lo = total * size_t(k) / size_t(nthreads), hi = total * size_t(k + 1) / size_t(nthreads);
    size_t pos = 0;
    for (const Seg& g : segs) {
      const size_t a = std::max(lo, pos), b = std::min(hi, pos + g.n);
      if (a < b)
        std::memcpy(g.dst + (a - pos), g.src + (a - pos), b - a);
      pos += g.n;
      if (pos >= hi)
        break;
    }
  };
  std::vector<std::thread> th;
  for (int k = 1; k < nthreads; k++)
    th.emplace_back(share, k);
  share(0);
  for (auto& t : th)
    t.join();
}
int copy_threads(int ndev)
{
  const int hw = int(std::thread::hardware_concurrency());
  return std::max(1, std::min(6, (hw > 2 ? hw - 2 : 1) / std::max(1, ndev)));
}

// host box -> device (linear box layout)
void box_h2d(MultiState& m, void* d_box, const BoxMap& b, int nthreads)
{
  size_t done = 0;
  std::vector<Seg> segs;
  for (size_t i = 0; done < b.bytes; i++) {
    const int s = int(i % kRing);
    const size_t len = std::min(kSlot, b.bytes - done);
    if (i >= size_t(kRing))
      RT_CHECK(cudaEventSynchronize(m.ev[s]));
    segs.clear();
    for (size_t o = 0; o < len;) {
      size_t left;
      const char* src = b.at(done + o, left);
      const size_t c = std::min(left, len - o);
      segs.push_back(Seg{static_cast<char*>(m.slot[s]) + o, src, c});
      o += c;
    }
    par_copy(segs, nthreads);
    RT_CHECK(cudaMemcpyAsync(static_cast<char*>(d_box) + done, m.slot[s], len, cudaMemcpyHostToDevice, m.st));
    RT_CHECK(cudaEventRecord(m.ev[s], m.st));
    done += len;
  }
  RT_CHECK(cudaStreamSynchronize(m.st));
}

// device (linear box layout) -> host box
void box_d2h(MultiState& m, const void* d_box, const BoxMap& b, int nthreads)
{
  const size_t n = (b.bytes + kSlot - 1) / kSlot;
  std::vector<Seg> segs;
  auto drain = [&](size_t i) {
    const int s = int(i % kRing);
    const size_t off = i * kSlot, len = std::min(kSlot, b.bytes - off);
    RT_CHECK(cudaEventSynchronize(m.ev[s]));
    segs.clear();
    for (size_t o = 0; o < len;) {
      size_t left;
      char* dst = b.at(off + o, left);
      const size_t c = std::min(left, len - o);
      segs.push_back(Seg{dst, static_cast<const char*>(m.slot[s]) + o, c});
      o += c;
    }
    par_copy(segs, nthreads);   // (the threads also fault the pages of the fresh result buffer in)
  };
  for (size_t i = 0; i < n; i++) {
    if (i >= size_t(kRing))
      drain(i - kRing);
    const int s = int(i % kRing);
    const size_t off = i * kSlot, len = std::min(kSlot, b.bytes - off);
    RT_CHECK(cudaMemcpyAsync(m.slot[s], static_cast<const char*>(d_box) + off, len, cudaMemcpyDeviceToHost, m.st));
    RT_CHECK(cudaEventRecord(m.ev[s], m.st));
  }
  for (size_t i = n > size_t(kRing) ? n - kRing : 0; i < n; i++)
    drain(i);
}

}  // namespace
#endif

namespace sperr_b200 {

// Returns -2 when the call cannot be split that way (the caller then takes the one-device path).
int comp_3d_multi(const void* src, int is_float, const size_t vol[3], const size_t chunk[3], int mode,
                  double quality, const std::vector<int>& devs_in, void** dst, size_t* dst_len)
{
#ifdef SPERR_EMUL
  return -2;
#else
  size_t cd[3];
  for (int i = 0; i < 3; i++)
    cd[i] = std::min(std::max<size_t>(1, chunk[i]), vol[i]);
  const size_t nchunks = sperr_b200_num_chunks(vol, cd);
  std::vector<int> devs = devs_in;
  std::vector<size_t> begins;
  while (devs.size() >= 2) {   // as many of the devices as give every one a box of whole chunks
    begins.assign(devs.size() + 1, 0);
    if (nchunks >= devs.size() && shard_ranges(vol, cd, devs.size(), begins.data()))
      break;
    devs.pop_back();
  }
  if (devs.size() < 2)
    return -2;
  const int nd = int(devs.size());
  const size_t esz = is_float ? 4 : 8;
  std::vector<uint32_t> lens(nchunks, 0);
  std::vector<size_t> nbytes(nd, 0), off(nd, 0);
  std::vector<int> rcs(nd, 0);
  std::atomic<int> failed{0};
  uint8_t* container = nullptr;
  size_t hlen = 0, total = 0;
  Barrier bar(nd);
  int home = 0;
  cudaGetDevice(&home);
  auto work = [&](int d) {
    const void* d_streams = nullptr;
    try {
      RT_CHECK(cudaSetDevice(devs[d]));
      MultiState& m = multi_state();
      m.init();
      size_t org[3], ext[3];
      if (sperr_b200_chunk_box(vol, cd, begins[d], begins[d + 1], org, ext) != 0)
        throw std::runtime_error("chunk box");
      const BoxMap bm(src, vol, org, ext, esz);
      m.box.reserve(bm.bytes);
      box_h2d(m, m.box.p, bm, copy_threads(nd));
      size_t n = 0;
      rcs[d] = sperr_b200_comp_3d_range_dev(m.box.p, is_float, vol, cd, org, ext, begins[d], begins[d + 1], mode,
                                            quality, &d_streams, &n, lens.data() + begins[d]);
      nbytes[d] = n;
      if (rcs[d] != 0)
        failed = 1;
    }
    catch (const std::exception& e) {
      if (std::getenv("SPERR_B200_VERBOSE"))
        std::fprintf(stderr, "sperr_b200 (device %d): %s\n", devs[d], e.what());
      rcs[d] = -1;
      failed = 1;
    }
    bar.wait();   // every length is known
    if (d == 0 && !failed) {
      hlen = sperr_b200_container_header(vol, cd, is_float, nullptr, nchunks, nullptr, 0);
      total = hlen;
      for (int k = 0; k < nd; k++) {
        off[k] = total;
        total += nbytes[k];
      }
      container = static_cast<uint8_t*>(std::malloc(total));
      if (!container)
        failed = 1;
      else
        sperr_b200_container_header(vol, cd, is_float, lens.data(), nchunks, container, hlen);
    }
    bar.wait();   // the container exists
    if (!failed && nbytes[d]) {
      if (cudaMemcpy(container + off[d], d_streams, nbytes[d], cudaMemcpyDeviceToHost) != cudaSuccess) {
        cudaGetLastError();
        failed = 1;
      }
    }
  };
  std::vector<std::thread> th;
  for (int d = 1; d < nd; d++)
    th.emplace_back(work, d);
  work(0);
  for (auto& t : th)
    t.join();
  cudaSetDevice(home);
  if (failed) {
    std::free(container);
    for (int d = 0; d < nd; d++)
      if (rcs[d] == 2)
        return 2;
    return -1;
  }
  *dst = container;
  *dst_len = total;
  return 0;
#endif
}

int decomp_3d_multi(const void* src, size_t src_len, int output_float, const std::vector<int>& devs_in,
                    size_t* dimx, size_t* dimy, size_t* dimz, void** dst)
{
#ifdef SPERR_EMUL
  return -2;
#else
  size_t vol[3], cd[3], hlen = 0, nchunks = 0;
  int is_float = 0;
  if (sperr_b200_parse_container(src, src_len, vol, cd, &is_float, &hlen, nullptr, 0, &nchunks) != 0)
    return -1;
  std::vector<uint32_t> lens(nchunks);
  if (sperr_b200_parse_container(src, src_len, vol, cd, &is_float, &hlen, lens.data(), nchunks, &nchunks) != 0)
    return -1;
  std::vector<size_t> coff(nchunks + 1, hlen);
  for (size_t i = 0; i < nchunks; i++)
    coff[i + 1] = coff[i] + lens[i];
  if (coff[nchunks] != src_len)
    return -1;
  std::vector<int> devs = devs_in;
  std::vector<size_t> begins;
  while (devs.size() >= 2) {
    begins.assign(devs.size() + 1, 0);
    if (nchunks >= devs.size() && shard_ranges(vol, cd, devs.size(), begins.data()))
      break;
    devs.pop_back();
  }
  if (devs.size() < 2)
    return -2;
  const int nd = int(devs.size());
  const size_t esz = output_float ? 4 : 8;
  const size_t total = vol[0] * vol[1] * vol[2];
  void* out = std::malloc(total * esz);
  if (!out)
    return -1;
  if (total * esz >= (size_t(64) << 20)) {
    const uintptr_t a = (reinterpret_cast<uintptr_t>(out) + 4095) & ~uintptr_t(4095);
    madvise(reinterpret_cast<void*>(a), (total * esz - (a - reinterpret_cast<uintptr_t>(out))) & ~size_t(4095),
            MADV_HUGEPAGE);
  }
  std::atomic<int> failed{0};
  int home = 0;
  cudaGetDevice(&home);
  auto work = [&](int d) {
    try {
      RT_CHECK(cudaSetDevice(devs[d]));
      MultiState& m = multi_state();
      m.init();
      size_t org[3], ext[3];
      if (sperr_b200_chunk_box(vol, cd, begins[d], begins[d + 1], org, ext) != 0)
        throw std::runtime_error("chunk box");
      const BoxMap bm(out, vol, org, ext, esz);
      m.box.reserve(bm.bytes);
      const size_t b0 = coff[begins[d]], b1 = coff[begins[d + 1]];
      const int rc = sperr_b200_decomp_3d_range_dev(static_cast<const uint8_t*>(src) + b0, nullptr, b1 - b0,
                                                    lens.data() + begins[d], vol, cd, org, ext, begins[d],
                                                    begins[d + 1], output_float, m.box.p);
      if (rc != 0)
        failed = 1;
      else
        box_d2h(m, m.box.p, bm, copy_threads(nd));
    }
    catch (const std::exception& e) {
      if (std::getenv("SPERR_B200_VERBOSE"))
        std::fprintf(stderr, "sperr_b200 (device %d): %s\n", devs[d], e.what());
      failed = 1;
    }
  };
  std::vector<std::thread> th;
  for (int d = 1; d < nd; d++)
    th.emplace_back(work, d);
  work(0);
  for (auto& t : th)
    t.join();
  cudaSetDevice(home);
  if (failed) {
    std::free(out);
    return -1;
  }
  *dimx = vol[0];
  *dimy = vol[1];
  *dimz = vol[2];
  *dst = out;
  return 0;
#endif
}

}  // namespace sperr_b200
