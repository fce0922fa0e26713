#include "hostpipe.h"

#include <algorithm>

#if defined(__x86_64__) && !defined(SPERR_EMUL)
#include <immintrin.h>
#define SPERR_B200_NT_COPY 1
#endif

namespace sperr_b200 {

namespace {

// Copy into a buffer that will not be read again soon (the caller's result array): non-temporal
// stores skip the read-for-ownership of every destination line, which is a third of the DRAM
// traffic of a plain memcpy. Source: pinned ring slot (cacheable).
#ifdef SPERR_B200_NT_COPY
__attribute__((target("avx2"))) void nt_copy_avx2(char* dst, const char* src, size_t n)
{
  const size_t head = (32 - (reinterpret_cast<uintptr_t>(dst) & 31)) & 31;
  if (head >= n) {
    std::memcpy(dst, src, n);
    return;
  }
  std::memcpy(dst, src, head);
  dst += head; src += head; n -= head;
  size_t i = 0;
  for (; i + 128 <= n; i += 128) {
    const __m256i a = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i));
    const __m256i b = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 32));
    const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 64));
    const __m256i d = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(src + i + 96));
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i), a);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 32), b);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 64), c);
    _mm256_stream_si256(reinterpret_cast<__m256i*>(dst + i + 96), d);
  }
  _mm_sfence();
  if (i < n)
    std::memcpy(dst + i, src + i, n - i);
}
#endif

void out_copy(char* dst, const char* src, size_t n)
{
#ifdef SPERR_B200_NT_COPY
  static const bool avx2 = __builtin_cpu_supports("avx2") && !std::getenv("SPERR_B200_NO_NT_COPY");
  if (avx2 && n >= 4096) {
    nt_copy_avx2(dst, src, n);
    return;
  }
#endif
  std::memcpy(dst, src, n);
}

}  // namespace

HostPipe& HostPipe::get()
{
  static HostPipe* p = new HostPipe();   // lives for the process: worker threads stay parked
  return *p;
}

HostPipe::HostPipe()
{
#ifndef SPERR_EMUL
  for (int i = 0; i < kSlots; i++) {
    slot_[i] = rt::hmalloc_pinned(kSlotBytes);
    RT_CHECK(cudaEventCreateWithFlags(&ev_[i], cudaEventDisableTiming));
  }
  const unsigned hw = std::thread::hardware_concurrency();
  unsigned n = std::min(16u, std::max(2u, hw > 3 ? hw - 2 : 2u));   // memcpy + first-touch faults scale
  if (const char* e = std::getenv("SPERR_B200_COPY_THREADS"))
    n = std::max(1, std::atoi(e));
  for (unsigned i = 0; i < n; i++)
    threads_.emplace_back([this] { worker(); });
#endif
}

HostPipe::~HostPipe()
{
  {
    std::lock_guard<std::mutex> l(mu_);
    stop_ = true;
  }
  cv_work_.notify_all();
  for (auto& t : threads_)
    t.join();
}

void HostPipe::worker()
{
  std::unique_lock<std::mutex> l(mu_);
  for (;;) {
    cv_work_.wait(l, [this] { return stop_ || !jobs_.empty(); });
    if (stop_)
      return;
    const Job j = jobs_.back();
    jobs_.pop_back();
    l.unlock();
    if (j.ctl)
      drain_worker(j.ctl);
    else if (j.src) {
      // pageable source -> pinned slot: non-temporal stores here too (the slot is read next by the DMA
      // engine, not by a core). Measured on B200 / 16 host cores, 4 GiB from pageable memory with the
      // upload overlapped with the coder: sperr_comp_3d 126 ms with plain memcpy, 117 - 120 ms with
      // these (SPERR_B200_H2D_NO_NT=1 switches back).
      static const bool nt = std::getenv("SPERR_B200_H2D_NO_NT") == nullptr;
      if (nt)
        out_copy(j.dst, j.src, j.len);
      else
        std::memcpy(j.dst, j.src, j.len);
    }
    else   // first-touch: fault the pages in
      for (size_t off = 0; off < j.len; off += 4096)   // a write access that changes nothing, so
        __atomic_fetch_or(j.dst + off, 0, __ATOMIC_RELAXED);   // it may race with the real copy
    l.lock();
    if (--((j.src || j.ctl) ? pending_ : pending_pf_) == 0)
      cv_done_.notify_all();
  }
}

void HostPipe::prefault_begin(void* dst, size_t bytes)
{
  if (threads_.empty() || bytes < (size_t(16) << 20))
    return;
  const size_t piece = size_t(8) << 20;
  {
    std::lock_guard<std::mutex> l(mu_);
    for (size_t off = 0; off < bytes; off += piece) {
      jobs_.insert(jobs_.begin(), {static_cast<char*>(dst) + off, nullptr, std::min(piece, bytes - off)});
      pending_pf_++;
    }
  }
  cv_work_.notify_all();
}

void HostPipe::wait_idle()
{
  std::unique_lock<std::mutex> l(mu_);
  cv_done_.wait(l, [this] { return pending_ == 0 && pending_pf_ == 0; });
}

void HostPipe::parallel_copy(void* dst, const void* src, size_t bytes)
{
  if (threads_.empty() || bytes < (size_t(1) << 20)) {
    std::memcpy(dst, src, bytes);
    return;
  }
  const size_t np = std::min<size_t>(threads_.size(), (bytes + (size_t(1) << 20) - 1) >> 20);
  const size_t piece = ((bytes + np - 1) / np + 4095) & ~size_t(4095);
  {
    std::lock_guard<std::mutex> l(mu_);
    for (size_t off = 0; off < bytes; off += piece) {
      jobs_.push_back({static_cast<char*>(dst) + off, static_cast<const char*>(src) + off,
                       std::min(piece, bytes - off)});
      pending_++;
    }
  }
  cv_work_.notify_all();
  std::unique_lock<std::mutex> l(mu_);
  cv_done_.wait(l, [this] { return pending_ == 0; });
}

#ifndef SPERR_EMUL
static bool is_pinned_host(const void* p)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost;
}
#endif

void HostPipe::d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st)
{
#ifdef SPERR_EMUL
  rt::d2h(dst_host, src_dev, bytes, st);
#else
  if (bytes < (size_t(4) << 20) || is_pinned_host(dst_host)) {
    rt::d2h(dst_host, src_dev, bytes, st);
    rt::sync(st);
    return;
  }
  const size_t n = (bytes + kSlotBytes - 1) / kSlotBytes;
  auto len_of = [&](size_t i) { return std::min(kSlotBytes, bytes - i * kSlotBytes); };
  if (threads_.empty()) {
    for (size_t i = 0; i < n; i++) {
      rt::d2h(slot_[0], static_cast<const char*>(src_dev) + i * kSlotBytes, len_of(i), st);
      rt::sync(st);
      std::memcpy(static_cast<char*>(dst_host) + i * kSlotBytes, slot_[0], len_of(i));
    }
    return;
  }
  static_assert(kSlots == 4, "StreamCtl::slot_done is sized for four slots");
  StreamCtl ctl;
  ctl.dst = static_cast<char*>(dst_host);
  ctl.bytes = bytes;
  ctl.nslots = n;
  for (int s = 0; s < kSlots; s++)
    ctl.slot_done[s].store(0);
  {
    std::lock_guard<std::mutex> l(mu_);
    for (size_t w = 0; w < threads_.size(); w++) {
      jobs_.push_back({nullptr, nullptr, 0, &ctl});
      pending_++;
    }
  }
  cv_work_.notify_all();
  // the DMA side: keep every free slot of the ring in flight, publish arrivals in order
  size_t issued = 0, synced = 0;
  try {
    while (synced < n) {
      while (issued < n && (issued < size_t(kSlots) ||
                            ctl.slot_done[issued % kSlots].load(std::memory_order_acquire) >=
                                long(kPieces) * long(issued / kSlots))) {
        const int s = int(issued % kSlots);
        rt::d2h(slot_[s], static_cast<const char*>(src_dev) + issued * kSlotBytes, len_of(issued), st);
        RT_CHECK(cudaEventRecord(ev_[s], st));
        issued++;
      }
      if (synced < issued) {
        RT_CHECK(cudaEventSynchronize(ev_[synced % kSlots]));
        synced++;
        ctl.dma_done.store(long(synced), std::memory_order_release);
      }
      else
        std::this_thread::yield();   // the ring is full of slots the workers are still draining
    }
  }
  catch (...) {
    ctl.dma_done.store(long(n) + (1l << 40), std::memory_order_release);   // let the workers run out
    std::unique_lock<std::mutex> l(mu_);
    cv_done_.wait(l, [this] { return pending_ == 0; });
    throw;
  }
  std::unique_lock<std::mutex> l(mu_);
  cv_done_.wait(l, [this] { return pending_ == 0; });
#endif
}

// Takes (slot, piece) items in order until none is left; an item waits for its slot's DMA only.
void HostPipe::drain_worker(StreamCtl* c)
{
  const long items = long(c->nslots) * kPieces;
  for (;;) {
    const long t = c->next.fetch_add(1, std::memory_order_relaxed);
    if (t >= items)
      return;
    const long i = t / kPieces, p = t % kPieces;
    while (c->dma_done.load(std::memory_order_acquire) <= i)
      std::this_thread::yield();
    const size_t len = std::min(kSlotBytes, c->bytes - size_t(i) * kSlotBytes);
    const size_t piece = ((len + kPieces - 1) / kPieces + 4095) & ~size_t(4095);
    const size_t a = std::min(len, size_t(p) * piece), b = std::min(len, a + piece);
    if (b > a && c->dma_done.load(std::memory_order_relaxed) < (1l << 40))
      out_copy(c->dst + size_t(i) * kSlotBytes + a, static_cast<const char*>(slot_[i % kSlots]) + a, b - a);
    c->slot_done[i % kSlots].fetch_add(1, std::memory_order_release);
    c->total_done.fetch_add(1, std::memory_order_relaxed);
  }
}

void HostPipe::h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st)
{
#ifdef SPERR_EMUL
  rt::h2d(dst_dev, src_host, bytes, st);
#else
  if (bytes < (size_t(4) << 20) || is_pinned_host(src_host)) {
    rt::h2d(dst_dev, src_host, bytes, st);
    rt::sync(st);
    return;
  }
  const size_t n = (bytes + kSlotBytes - 1) / kSlotBytes;
  for (size_t i = 0; i < n; i++) {
    const int s = int(i % kSlots);
    const size_t len = std::min(kSlotBytes, bytes - i * kSlotBytes);
    if (i >= size_t(kSlots))
      RT_CHECK(cudaEventSynchronize(ev_[s]));   // the DMA that last read this slot has finished
    parallel_copy(slot_[s], static_cast<const char*>(src_host) + i * kSlotBytes, len);
    rt::h2d(static_cast<char*>(dst_dev) + i * kSlotBytes, slot_[s], len, st);
    RT_CHECK(cudaEventRecord(ev_[s], st));
  }
  rt::sync(st);
#endif
}

}  // namespace sperr_b200
