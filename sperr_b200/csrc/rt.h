// Thin runtime layer: CUDA under nvcc; under -DSPERR_EMUL (tests/emul, test infrastructure only)
// the same kernels are executed by a CPU SIMT emulator so their logic can be debugged in a
// container without a GPU. The product library is always the nvcc build.
#pragma once

#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <chrono>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#ifdef SPERR_EMUL
#include "cuda_emul.h"
typedef int cudaStream_t;
#define LAUNCH(kern, grid, block, smem, stream, ...) \
  emu::launch(grid, block, smem, [=]() { kern(__VA_ARGS__); })
#define LAUNCH_CLUSTER(kern, grid, block, smem, stream, csize, ...) \
  emu::launch_cluster(grid, block, smem, csize, [=]() { kern(__VA_ARGS__); })
#define DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem())
#define HD
#define ASSUME_GLOBAL(p) ((void)0)
#else
#include <cuda_runtime.h>
#define LAUNCH(kern, grid, block, smem, stream, ...)        \
  do {                                                      \
    kern<<<grid, block, smem, stream>>>(__VA_ARGS__);       \
    rt::check(cudaGetLastError(), #kern, __FILE__, __LINE__); \
    rt::launch_counter()++;                                 \
  } while (0)
// thread-block clusters of `csize` consecutive CTAs (1D grids)
#define LAUNCH_CLUSTER(kern, grid_, block_, smem_, stream_, csize, ...)                      \
  do {                                                                                       \
    cudaLaunchConfig_t cfg_ = {};                                                            \
    cfg_.gridDim = grid_;                                                                    \
    cfg_.blockDim = block_;                                                                  \
    cfg_.dynamicSmemBytes = smem_;                                                           \
    cfg_.stream = stream_;                                                                   \
    cudaLaunchAttribute at_[1];                                                              \
    at_[0].id = cudaLaunchAttributeClusterDimension;                                         \
    at_[0].val.clusterDim.x = csize;                                                         \
    at_[0].val.clusterDim.y = 1;                                                             \
    at_[0].val.clusterDim.z = 1;                                                             \
    cfg_.attrs = at_;                                                                        \
    cfg_.numAttrs = 1;                                                                       \
    rt::check(cudaLaunchKernelEx(&cfg_, kern, __VA_ARGS__), #kern, __FILE__, __LINE__);      \
    rt::launch_counter()++;                                                                  \
  } while (0)
#define DYN_SMEM(type, name)                                        \
  extern __shared__ __align__(16) unsigned char name##_raw_smem[];  \
  type* name = reinterpret_cast<type*>(name##_raw_smem)
#define HD __host__ __device__
// A pointer fetched from a descriptor in memory is a generic pointer to the compiler: it emits LD /
// ST, which may alias shared memory and therefore pin every shared-memory access behind them in
// program order. This tells it the pointer is a global one (LDG / STG).
#define ASSUME_GLOBAL(p) __builtin_assume(__isGlobal(p))
#endif

// The same hint as an expression: gptr(ch.coef)[i]
template <class T>
__device__ __forceinline__ T* gptr(T* p)
{
  ASSUME_GLOBAL(p);
  return p;
}

// ---- CTA barrier that does not require the threads of a warp to arrive together. What
// __syncthreads() compiles to is barrier.sync.aligned: every thread of a warp has to execute the same
// barrier instruction at the same time. Kernels whose warps run long single-thread sections between
// barriers (thread 0 walking a tree, thread 0 spinning on a grid barrier) do not always have their
// warps reconverged at the barrier after optimisation; on B200 that showed as decoder state that was
// one barrier out of step (compute-sanitizer synccheck: "divergent thread(s) in warp"). ----
#ifdef SPERR_EMUL
inline void cta_sync() { __syncthreads(); }
#elif defined(__CUDACC__)
__device__ __forceinline__ void cta_sync()
{
  asm volatile("barrier.sync 0;" ::: "memory");
}
#endif

// ---- thread-block clusters: rank of the CTA and the cluster-wide barrier (release / acquire, so
// what a CTA wrote to global memory before it is visible to the others after it) ----
#ifdef SPERR_EMUL
inline unsigned cluster_rank() { return emu::cluster_rank(); }
inline void cluster_sync() { emu::sync_cluster(); }
#elif defined(__CUDACC__)
__device__ __forceinline__ unsigned cluster_rank()
{
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync()
{
  // NOT the .aligned form: that one requires every thread of a warp to execute the same barrier
  // instruction together, which the decoders' control flow (long single-thread parts between
  // barriers) does not guarantee after optimisation -- measured on B200: with .aligned the cluster
  // decode of a 128^3 chunk hangs, without it it is bit-exact.
  __threadfence();
  asm volatile("barrier.cluster.arrive.release;\n\tbarrier.cluster.wait.acquire;" ::: "memory");
  __syncwarp();   // ... and leave converged: the CTA barriers that follow are the aligned kind
}
#endif

namespace rt {

// The CUDA device of the calling host thread (0 under the emulator). Everything the library keeps
// between calls -- work buffers, helper streams, per-kernel attributes -- is kept per device, so
// that host threads driving different GPUs never share state.
constexpr int kMaxDevices = 16;
inline int cur_dev()
{
#ifndef SPERR_EMUL
  int d = 0;
  if (cudaGetDevice(&d) != cudaSuccess)
    d = 0;
  return d < kMaxDevices ? d : 0;
#else
  return 0;
#endif
}
// true exactly once per device: guards one-time per-device set-up (cudaFuncSetAttribute, ...)
struct OncePerDevice {
  bool done[kMaxDevices] = {};
  bool first()
  {
    const int d = cur_dev();
    if (done[d])
      return false;
    done[d] = true;
    return true;
  }
};

// number of kernels of this library launched so far (bench.py reports the count per step)
inline std::atomic<unsigned long long>& launch_counter()
{
  static std::atomic<unsigned long long> n{0};
  return n;
}

#ifndef SPERR_EMUL
inline void check(cudaError_t e, const char* what, const char* file, int line)
{
  if (e != cudaSuccess) {
    std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " at " + what + " (" +
                      file + ":" + std::to_string(line) + ")";
    throw std::runtime_error(msg);
  }
}
#define RT_CHECK(x) rt::check((x), #x, __FILE__, __LINE__)

inline void* dmalloc(size_t n)
{
  void* p = nullptr;
  RT_CHECK(cudaMalloc(&p, n ? n : 1));
  return p;
}
inline void dfree(void* p)
{
  if (p)
    cudaFree(p);
}
inline void h2d(void* d, const void* s, size_t n, cudaStream_t st)
{
  if (n)
    RT_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyHostToDevice, st));
}
inline void d2d(void* d, const void* s, size_t n, cudaStream_t st)
{
  if (n)
    RT_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToDevice, st));
}
inline void dset(void* d, int v, size_t n, cudaStream_t st)
{
  if (n)
    RT_CHECK(cudaMemsetAsync(d, v, n, st));
}
inline void* hmalloc_pinned(size_t n)
{
  void* p = nullptr;
  RT_CHECK(cudaMallocHost(&p, n ? n : 1));
  return p;
}
inline void hfree_pinned(void* p)
{
  if (p)
    cudaFreeHost(p);
}
inline bool is_pinned_host(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}

// Small read-backs (counters, chunk descriptors) into ordinary host memory. A cudaMemcpyAsync into
// pageable memory only returns when the copy has been done, i.e. after everything queued on the
// stream before it, and the other host threads of the process cannot launch meanwhile: the outlier
// thread of the compressor, reading a counter behind a 13 ms kernel, stalled the SPECK encoder's
// thread for as long. Such copies land in a pinned bounce buffer of the stream instead (really
// asynchronous) and are handed to their destinations by the rt::sync of that stream. A stream is
// used by one host thread at a time, every d2h is followed by the sync of its stream on the same
// thread; copies into pinned memory and large ones are issued directly, as before. Opt-in per host
// thread (ReadbackScope): the compressor's two threads use it, everything else copies as before.
constexpr size_t kReadbackMax = size_t(1) << 20;
inline bool& readback_on()
{
  static thread_local bool on = false;
  return on;
}
struct ReadbackScope {
  bool prev;
  ReadbackScope() : prev(readback_on()) { readback_on() = true; }
  ~ReadbackScope() { readback_on() = prev; }
};
struct Readback {
  struct Item {
    void* dst;
    size_t off, n;
  };
  char* pin = nullptr;
  size_t cap = 0, used = 0;
  std::vector<Item> items;
};
struct ReadbackTable {
  std::mutex mu;
  // keyed by (device, stream): the legacy stream is handle 0 on EVERY device, and the threads that
  // drive different devices (capi_multi.cu) use it at the same time
  std::map<std::pair<int, cudaStream_t>, Readback> m;
};
inline ReadbackTable& readback_table()
{
  static ReadbackTable* t = new ReadbackTable();   // lives as long as the process
  return *t;
}
inline Readback& readback_of(cudaStream_t st)
{
  ReadbackTable& t = readback_table();
  const int dev = cur_dev();
  std::lock_guard<std::mutex> l(t.mu);
  return t.m[std::make_pair(dev, st)];
}
// Forgets copies whose rt::sync never came (a call that ended in an exception between the two):
// their destinations are gone. Only for a point where no read-back of the process is in flight,
// i.e. the start of an API call (callers are serialised) before it starts helper threads.
inline void readback_abandon()
{
  ReadbackTable& t = readback_table();
  const int dev = cur_dev();
  std::lock_guard<std::mutex> l(t.mu);
  for (auto& kv : t.m) {
    if (kv.first.first != dev)
      continue;   // another device's thread may have copies in flight (calls are serialised per device)
    kv.second.items.clear();
    kv.second.used = 0;
  }
}
inline void d2h(void* d, const void* s, size_t n, cudaStream_t st)
{
  if (!n)
    return;
  if (readback_on() && n <= kReadbackMax && !is_pinned_host(d)) {
    Readback& r = readback_of(st);
    const size_t need = (n + 15) & ~size_t(15);
    if (r.used + need > r.cap && r.items.empty()) {   // nothing in flight: the buffer may be replaced
      hfree_pinned(r.pin);
      r.pin = nullptr;
      r.cap = 0;
      r.pin = static_cast<char*>(hmalloc_pinned(4 * kReadbackMax));
      r.cap = 4 * kReadbackMax;
      r.used = 0;
    }
    if (r.used + need <= r.cap) {
      RT_CHECK(cudaMemcpyAsync(r.pin + r.used, s, n, cudaMemcpyDeviceToHost, st));
      r.items.push_back(Readback::Item{d, r.used, n});
      r.used += need;
      return;
    }
  }
  RT_CHECK(cudaMemcpyAsync(d, s, n, cudaMemcpyDeviceToHost, st));
}
inline void sync(cudaStream_t st)
{
  RT_CHECK(cudaStreamSynchronize(st));
  Readback& r = readback_of(st);
  for (const Readback::Item& it : r.items)
    std::memcpy(it.dst, r.pin + it.off, it.n);
  r.items.clear();
  r.used = 0;
}
inline bool is_device_ptr(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}
#else
inline void* dmalloc(size_t n) { return std::malloc(n ? n : 1); }
inline void dfree(void* p) { std::free(p); }
inline void h2d(void* d, const void* s, size_t n, cudaStream_t) { std::memcpy(d, s, n); }
inline void d2h(void* d, const void* s, size_t n, cudaStream_t) { std::memcpy(d, s, n); }
inline void d2d(void* d, const void* s, size_t n, cudaStream_t) { std::memmove(d, s, n); }
inline void dset(void* d, int v, size_t n, cudaStream_t) { std::memset(d, v, n); }
inline void sync(cudaStream_t) {}
struct ReadbackScope {};
inline void readback_abandon() {}
inline void* hmalloc_pinned(size_t n) { return std::malloc(n ? n : 1); }
inline void hfree_pinned(void* p) { std::free(p); }
inline bool is_device_ptr(const void*) { return false; }
#endif

// ---- stage profiler: CUDA-event ranges on the launching stream, off unless enabled ----
struct ProfEntry {
  const char* name;
#ifndef SPERR_EMUL
  cudaEvent_t e0, e1;
  int dev;
#endif
};
struct ProfState {
  bool on = false;
  std::mutex mu;   // ranges may be recorded by the pipelines' helper threads
  std::vector<ProfEntry> open;                       // recorded, not yet resolved
  std::map<std::string, std::pair<double, long>> acc;  // name -> (ms, ranges)
  // host wall time between the two records of a range: (sum, longest) in ms -- a host-side stall
  // (a blocked API call, a descheduled thread) shows here and not in the device times
  std::map<std::string, std::pair<double, double>> host;
#ifndef SPERR_EMUL
  // Events are REUSED. Creating two per range inside a timed loop (the first version) stalled the
  // host now and then for 15 - 110 ms -- measured on B200: steps of 119 ms became 130 - 230 ms one
  // time in three with the profiler on, never with it off -- presumably when the driver has to grow
  // its event storage under a busy GPU. The pool is filled when the profiler is switched on and
  // finished ranges hand their events back as soon as a later range finds them complete.
  std::map<int, std::vector<cudaEvent_t>> pool;      // per device
#endif
};
inline ProfState& prof()
{
  static ProfState s;
  return s;
}
#ifndef SPERR_EMUL
// diagnosis: SPERR_B200_PROF_NOTIMING=1 records events without time stamps (all ranges read 0 ms)
inline unsigned prof_event_flags()
{
  static const bool notiming = std::getenv("SPERR_B200_PROF_NOTIMING") != nullptr;
  return notiming ? cudaEventDisableTiming : cudaEventDefault;
}
inline cudaEvent_t prof_event_locked(int dev)
{
  auto& v = prof().pool[dev];
  if (!v.empty()) {
    cudaEvent_t e = v.back();
    v.pop_back();
    return e;
  }
  cudaEvent_t e;
  cudaEventCreateWithFlags(&e, prof_event_flags());
  return e;
}
// resolves the ranges at the front of the list that have completed (never waits)
inline void prof_resolve_done_locked(bool wait)
{
  auto& open = prof().open;
  size_t done = 0;
  for (; done < open.size(); done++) {
    ProfEntry& e = open[done];
    if (wait)
      cudaEventSynchronize(e.e1);
    else if (cudaEventQuery(e.e1) != cudaSuccess) {
      cudaGetLastError();
      break;
    }
    float ms = 0;
    if (cudaEventElapsedTime(&ms, e.e0, e.e1) != cudaSuccess) {
      cudaGetLastError();
      if (!wait)
        break;
    }
    auto& a = prof().acc[e.name];
    a.first += ms;
    a.second += 1;
    auto& v = prof().pool[e.dev];
    v.push_back(e.e0);
    v.push_back(e.e1);
  }
  open.erase(open.begin(), open.begin() + done);
}
// fills the current device's pool (called when the profiler is switched on, outside any timed region)
inline void prof_reserve(size_t events)
{
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> l(prof().mu);
  auto& v = prof().pool[dev];
  while (v.size() < events) {
    cudaEvent_t e;
    if (cudaEventCreateWithFlags(&e, prof_event_flags()) != cudaSuccess) {
      cudaGetLastError();
      break;
    }
    v.push_back(e);
  }
}
#endif
struct ProfScope {
#ifndef SPERR_EMUL
  cudaStream_t st;
  ProfEntry e;
  bool live;
  std::chrono::steady_clock::time_point t0;
  // diagnosis: SPERR_B200_PROF_ONLY=<prefix> records only the ranges whose name starts with it
  static bool wanted(const char* name)
  {
    static const char* only = std::getenv("SPERR_B200_PROF_ONLY");
    return !only || std::strncmp(name, only, std::strlen(only)) == 0;
  }
  ProfScope(const char* name, cudaStream_t s) : st(s), live(prof().on && wanted(name))
  {
    if (!live)
      return;
    e.name = name;
    cudaGetDevice(&e.dev);
    {
      std::lock_guard<std::mutex> l(prof().mu);
      e.e0 = prof_event_locked(e.dev);
      e.e1 = prof_event_locked(e.dev);
    }
    cudaEventRecord(e.e0, st);
    t0 = std::chrono::steady_clock::now();
  }
  ~ProfScope()
  {
    if (!live)
      return;
    cudaEventRecord(e.e1, st);
    const double hms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
    std::lock_guard<std::mutex> l(prof().mu);
    auto& h = prof().host[e.name];
    h.first += hms;
    h.second = std::max(h.second, hms);
    prof().open.push_back(e);
    if (prof().open.size() >= 256)
      prof_resolve_done_locked(false);
  }
#else
  ProfScope(const char*, cudaStream_t) {}
#endif
};
// Resolves all recorded ranges (synchronises) and adds them to the accumulators.
inline void prof_collect()
{
#ifndef SPERR_EMUL
  std::lock_guard<std::mutex> l(prof().mu);
  prof_resolve_done_locked(true);
#endif
  prof().open.clear();
}

// RAII device buffer
struct DBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DBuf() = default;
  explicit DBuf(size_t n) { alloc(n); }
  DBuf(const DBuf&) = delete;
  DBuf& operator=(const DBuf&) = delete;
  DBuf(DBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; o.bytes = 0; }
  DBuf& operator=(DBuf&& o) noexcept
  {
    if (this != &o) {
      release();
      p = o.p; bytes = o.bytes;
      o.p = nullptr; o.bytes = 0;
    }
    return *this;
  }
  ~DBuf() { release(); }
  void alloc(size_t n)
  {
    release();
    p = dmalloc(n);
    bytes = n;
    held() += n;
  }
  // device bytes all DBufs of this process hold right now (grow-only work buffers are reused by the
  // next call, so batch sizing counts them as available)
  static std::atomic<size_t>& held()
  {
    static std::atomic<size_t> h{0};
    return h;
  }
  // grow-only
  void reserve(size_t n)
  {
    if (n > bytes)
      alloc(n);
  }
  void release()
  {
    dfree(p);
    if (p)
      held() -= bytes;
    p = nullptr;
    bytes = 0;
  }
  template <typename T>
  T* as() const
  {
    return reinterpret_cast<T*>(p);
  }
};

}  // namespace rt
