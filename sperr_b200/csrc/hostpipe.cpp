#include "hostpipe.h"

#include <algorithm>

namespace sperr_b200 {

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
    if (j.src)
      std::memcpy(j.dst, j.src, j.len);
    else   // first-touch: fault the pages in
      for (size_t off = 0; off < j.len; off += 4096)   // a write access that changes nothing, so
        __atomic_fetch_or(j.dst + off, 0, __ATOMIC_RELAXED);   // it may race with the real copy
    l.lock();
    if (--(j.src ? pending_ : pending_pf_) == 0)
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
  for (size_t i = 0; i < n + kSlots - 1; i++) {
    if (i < n) {
      const int s = int(i % kSlots);
      rt::d2h(slot_[s], static_cast<const char*>(src_dev) + i * kSlotBytes, len_of(i), st);
      RT_CHECK(cudaEventRecord(ev_[s], st));
    }
    if (i + 1 >= size_t(kSlots)) {
      const size_t k = i + 1 - kSlots;
      if (k < n) {
        const int s = int(k % kSlots);
        RT_CHECK(cudaEventSynchronize(ev_[s]));
        parallel_copy(static_cast<char*>(dst_host) + k * kSlotBytes, slot_[s], len_of(k));
      }
    }
  }
#endif
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
