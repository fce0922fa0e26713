// TEST INFRASTRUCTURE ONLY -- runtime of the CPU SIMT emulator (see cuda_emul.h).
#include "cuda_emul.h"

#include <pthread.h>
#include <sys/mman.h>

#include <atomic>
#include <thread>
#include <vector>

extern "C" void emu_switch(void** from_sp, void* to_sp);
asm(R"(
.text
.globl emu_switch
.type emu_switch,@function
emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size emu_switch,.-emu_switch
)");

namespace emu {

enum State { RUNNABLE, WAIT_BLOCK, WAIT_CLUSTER, DONE };

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  uint3 tid{0, 0, 0};
  unsigned linear = 0;
  State st = DONE;
};

struct Warp {
  uint64_t slot[32];
  int arrived = 0, departed = 0;
  bool reading = false;
  int alive = 32;
};

static const size_t kStack = 256 * 1024;

struct Worker {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  void* sched_sp = nullptr;
  Fiber* current = nullptr;
  uint3 bid{0, 0, 0};
  dim3 bdim, gdim;
  std::vector<char> smem;
  const std::function<void()>* body = nullptr;
  // thread-block clusters: rank inside the cluster and the barrier its workers share
  unsigned crank = 0, csize = 1;
  pthread_barrier_t* cbar = nullptr;
};

static thread_local Worker* W = nullptr;
static thread_local uint3 host_tid{0, 0, 0}, host_bid{0, 0, 0};
static thread_local dim3 host_bdim, host_gdim;

Fiber* cur() { return W ? W->current : nullptr; }
uint3& tid() { return W && W->current ? W->current->tid : host_tid; }
uint3& bid() { return W ? W->bid : host_bid; }
dim3& bdim() { return W ? W->bdim : host_bdim; }
dim3& gdim() { return W ? W->gdim : host_gdim; }
unsigned lane() { return W->current->linear & 31; }
void* dyn_smem() { return W->smem.data(); }

static void yield_to_sched()
{
  Fiber* f = W->current;
  emu_switch(&f->sp, W->sched_sp);
}

unsigned cluster_rank() { return W ? W->crank : 0; }
unsigned cluster_size() { return W ? W->csize : 1; }
void sync_cluster()
{
  W->current->st = WAIT_CLUSTER;
  yield_to_sched();
}

void sync_block()
{
  W->current->st = WAIT_BLOCK;
  yield_to_sched();
}

void warp_exchange(uint64_t mine, uint64_t* out)
{
  Fiber* f = W->current;
  Warp& w = W->warps[f->linear >> 5];
  if (w.alive != 32) {
    std::fprintf(stderr, "emu: warp collective with exited/partial lanes (block %u,%u thread %u)\n",
                 W->bid.x, W->bid.y, f->linear);
    std::abort();
  }
  while (w.reading)
    yield_to_sched();
  w.slot[f->linear & 31] = mine;
  if (++w.arrived == 32)
    w.reading = true;
  else
    while (!w.reading)
      yield_to_sched();
  std::memcpy(out, w.slot, sizeof(w.slot));
  if (++w.departed == 32) {
    w.arrived = w.departed = 0;
    w.reading = false;
  }
}

static void fiber_main()
{
  (*W->body)();
  Fiber* f = W->current;
  f->st = DONE;
  W->warps[f->linear >> 5].alive--;
  for (;;)
    yield_to_sched();
}

static void prep_fiber(Fiber& f)
{
  if (!f.stack) {
    f.stack = (char*)mmap(nullptr, kStack, PROT_READ | PROT_WRITE,
                          MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
    if (f.stack == MAP_FAILED) {
      std::perror("mmap");
      std::abort();
    }
  }
  uintptr_t top = (uintptr_t)(f.stack + kStack);
  top &= ~uintptr_t(15);
  void** s = (void**)top;
  *--s = nullptr;                // fake return address of fiber_main
  *--s = (void*)&fiber_main;     // popped by `ret` in emu_switch
  for (int i = 0; i < 6; i++)
    *--s = nullptr;              // rbp rbx r12..r15
  f.sp = (void*)s;
}

static void run_block(Worker& wk, unsigned nthreads)
{
  const unsigned nwarps = nthreads / 32;
  for (unsigned t = 0; t < nthreads; t++) {
    Fiber& f = wk.fibers[t];
    prep_fiber(f);
    f.linear = t;
    f.tid.x = t % wk.bdim.x;
    f.tid.y = (t / wk.bdim.x) % wk.bdim.y;
    f.tid.z = t / (wk.bdim.x * wk.bdim.y);
    f.st = RUNNABLE;
  }
  for (unsigned w = 0; w < nwarps; w++)
    wk.warps[w] = Warp();
  for (;;) {
    unsigned done = 0, waiting = 0;
    for (unsigned w = 0; w < nwarps; w++) {
      // run this warp until every lane is at a block barrier or finished
      for (;;) {
        bool any = false;
        for (unsigned l = 0; l < 32; l++) {
          Fiber& f = wk.fibers[w * 32 + l];
          if (f.st == RUNNABLE) {
            any = true;
            wk.current = &f;
            emu_switch(&wk.sched_sp, f.sp);
            wk.current = nullptr;
          }
        }
        if (!any)
          break;
      }
      for (unsigned l = 0; l < 32; l++) {
        State s = wk.fibers[w * 32 + l].st;
        done += (s == DONE);
        waiting += (s == WAIT_BLOCK || s == WAIT_CLUSTER);
      }
    }
    if (done == nthreads)
      break;
    // every live thread is at the barrier: release
    (void)waiting;
    bool at_cluster = false;
    for (unsigned t = 0; t < nthreads; t++)
      at_cluster |= wk.fibers[t].st == WAIT_CLUSTER;
    if (at_cluster) {
      for (unsigned t = 0; t < nthreads; t++)
        if (wk.fibers[t].st == WAIT_BLOCK) {
          std::fprintf(stderr, "emu: block %u mixes __syncthreads and cluster barriers\n", wk.bid.x);
          std::abort();
        }
      if (wk.cbar)
        pthread_barrier_wait(wk.cbar);   // the other blocks of the cluster arrive the same way
    }
    for (unsigned t = 0; t < nthreads; t++)
      if (wk.fibers[t].st == WAIT_BLOCK || wk.fibers[t].st == WAIT_CLUSTER)
        wk.fibers[t].st = RUNNABLE;
  }
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body)
{
  const unsigned nthreads = block.x * block.y * block.z;
  if (nthreads == 0 || nthreads % 32 != 0 || nthreads > 1024) {
    std::fprintf(stderr, "emu: block size %u must be a multiple of 32 and <= 1024\n", nthreads);
    std::abort();
  }
  const uint64_t nblocks = uint64_t(grid.x) * grid.y * grid.z;
  if (nblocks == 0)
    return;
  static std::vector<Worker*> pool;  // launches are issued from one host thread at a time
  unsigned nworkers = (unsigned)std::min<uint64_t>(nblocks, 8);
  if (const char* e = std::getenv("EMU_WORKERS"))
    nworkers = (unsigned)std::max(1, std::min(atoi(e), (int)nworkers));
  while (pool.size() < nworkers)
    pool.push_back(new Worker());
  std::atomic<uint64_t> next{0};
  auto work = [&](unsigned wi) {
    Worker& wk = *pool[wi];
    W = &wk;
    if (wk.fibers.size() < nthreads)
      wk.fibers.resize(nthreads);
    if (wk.warps.size() < nthreads / 32)
      wk.warps.resize(nthreads / 32);
    if (wk.smem.size() < smem + 64)
      wk.smem.resize(smem + 64);
    wk.bdim = block;
    wk.gdim = grid;
    wk.body = &body;
    for (;;) {
      uint64_t b = next.fetch_add(1);
      if (b >= nblocks)
        break;
      wk.bid.x = unsigned(b % grid.x);
      wk.bid.y = unsigned((b / grid.x) % grid.y);
      wk.bid.z = unsigned(b / (uint64_t(grid.x) * grid.y));
      run_block(wk, nthreads);
    }
    W = nullptr;
  };
  if (nworkers == 1)
    work(0);
  else {
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nworkers; i++)
      th.emplace_back(work, i);
    for (auto& t : th)
      t.join();
  }
}

// Clusters of `csize` consecutive blocks (1D grids): the blocks of a cluster run at the same time on
// csize OS threads that share a barrier. Every block of a cluster must reach every cluster barrier.
void launch_cluster(dim3 grid, dim3 block, size_t smem, unsigned csize, const std::function<void()>& body)
{
  const unsigned nthreads = block.x * block.y * block.z;
  if (csize <= 1) {
    launch(grid, block, smem, body);
    return;
  }
  if (grid.y != 1 || grid.z != 1 || grid.x % csize != 0 || nthreads % 32 != 0 || nthreads > 1024) {
    std::fprintf(stderr, "emu: bad cluster launch\n");
    std::abort();
  }
  const unsigned nclusters = grid.x / csize;
  const unsigned ngroups = std::max(1u, std::min(nclusters, 8u / csize));
  std::vector<Worker*> ws(size_t(ngroups) * csize);
  for (auto& w : ws)
    w = new Worker();
  std::vector<pthread_barrier_t> bars(ngroups);
  for (auto& b : bars)
    pthread_barrier_init(&b, nullptr, csize);
  auto work = [&](unsigned g, unsigned r) {
    Worker& wk = *ws[size_t(g) * csize + r];
    W = &wk;
    wk.fibers.resize(nthreads);
    wk.warps.resize(nthreads / 32);
    wk.smem.resize(smem + 64);
    wk.bdim = block;
    wk.gdim = grid;
    wk.body = &body;
    wk.crank = r;
    wk.csize = csize;
    wk.cbar = &bars[g];
    for (unsigned c = g; c < nclusters; c += ngroups) {
      wk.bid.x = c * csize + r;
      wk.bid.y = wk.bid.z = 0;
      run_block(wk, nthreads);
      pthread_barrier_wait(wk.cbar);   // the next cluster starts when every block of this one is done
    }
    W = nullptr;
  };
  std::vector<std::thread> th;
  for (unsigned g = 0; g < ngroups; g++)
    for (unsigned r = 0; r < csize; r++)
      th.emplace_back(work, g, r);
  for (auto& t : th)
    t.join();
  for (auto& b : bars)
    pthread_barrier_destroy(&b);
  for (auto w : ws) {
    for (auto& f : w->fibers)
      if (f.stack)
        munmap(f.stack, kStack);
    delete w;
  }
}

}  // namespace emu
