// Host <-> device staging for the host-pointer C API.
//
// The reference API hands us pageable user memory and expects malloc'd results
// (/root/reference/include/SPERR_C_API.h:21-31). cudaMemcpy on such memory is staged by the driver
// at a few GB/s; here the copies go through a small ring of pinned slots instead, with a pool of
// host threads doing the pageable side (memcpy, including the first-touch page faults of a fresh
// malloc) while the DMA engine moves the neighbouring slot. Pinned user memory is copied directly.
#pragma once

#include <atomic>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "rt.h"

namespace sperr_b200 {

class HostPipe {
 public:
  static HostPipe& get();
  // dst_host may be pageable. Returns after the last byte is in dst_host.
  void d2h(void* dst_host, const void* src_dev, size_t bytes, cudaStream_t st);
  // Returns after the last byte has been handed to the DMA engine AND the copy has completed.
  void h2d(void* dst_dev, const void* src_host, size_t bytes, cudaStream_t st);
  // Starts touching every page of a freshly malloc'd buffer on the worker threads (returns at once),
  // so that the page faults overlap GPU work; wait_idle() before the buffer is filled.
  void prefault_begin(void* dst, size_t bytes);
  void wait_idle();
  // Copies [0, bytes) with all workers; used for the pageable side.
  void parallel_copy(void* dst, const void* src, size_t bytes);

 private:
  HostPipe();
  ~HostPipe();
  void worker();

  static constexpr size_t kSlotBytes = size_t(32) << 20;
  static constexpr int kSlots = 4;
  void* slot_[kSlots] = {nullptr, nullptr, nullptr, nullptr};
#ifndef SPERR_EMUL
  cudaEvent_t ev_[kSlots];
#endif
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_work_, cv_done_;
  // Device -> pageable host as one pipeline: the caller keeps the DMA engine busy on the ring of
  // pinned slots, the workers drain the slots piece by piece as they arrive (no fork / join per slot).
  static constexpr int kPieces = 16;   // pieces of a slot handed to the workers
  struct StreamCtl {
    char* dst = nullptr;
    size_t bytes = 0, nslots = 0;
    std::atomic<long> dma_done{0};       // slots [0, dma_done) have landed in the ring
    std::atomic<long> next{0};           // next (slot, piece) item to take
    std::atomic<long> slot_done[4];      // pieces copied out of ring slot s, over all its occupants
    std::atomic<long> total_done{0};
  };
  void drain_worker(StreamCtl* c);
  struct Job {
    char* dst;
    const char* src;
    size_t len;
    StreamCtl* ctl = nullptr;
  };
  std::vector<Job> jobs_;
  size_t pending_ = 0, pending_pf_ = 0;   // copy jobs / page-touch jobs in flight
  bool stop_ = false;
};

}  // namespace sperr_b200
