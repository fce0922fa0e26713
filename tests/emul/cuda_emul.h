// TEST INFRASTRUCTURE ONLY.
//
// A tiny CPU "SIMT emulator" so that the CUDA sources under sperr_b200/csrc can be compiled with
// g++ and executed, thread by thread, in the build container (which has no GPU). It exists to
// debug kernel logic against the oracle before spending GPU minutes; it is never linked into the
// product library, and the product has no CPU execution path.
//
// Model: every CUDA thread of a block is a ucontext fiber; fibers of one block run round-robin on
// one OS thread and switch at __syncthreads / warp collectives; blocks are distributed over a few
// OS worker threads. __shared__ becomes `static thread_local` (one copy per worker = per running
// block). Restrictions (our kernels obey them): blockDim.x*y*z is a multiple of 32; warp
// collectives are called by all 32 lanes with a full mask; no inter-block waiting.
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>

#define SPERR_EMUL 1

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __noinline__
#define __launch_bounds__(...)
#define __shared__ static thread_local
#define __align__(n) alignas(n)

struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
struct double2 {
  double x, y;
};
struct float4 {
  float x, y, z, w;
};
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
struct float2 {
  float x, y;
};
struct uint2 {
  unsigned x, y;
};
struct uint4 {
  unsigned x, y, z, w;
};
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

namespace emu {
struct Fiber;
Fiber* cur();
uint3& tid();
uint3& bid();
dim3& bdim();
dim3& gdim();
void sync_block();
// Exchanges one 64-bit value per lane across the caller's warp; out[32] receives all lanes.
void warp_exchange(uint64_t mine, uint64_t* out);
unsigned lane();
void* dyn_smem();
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body);
// thread-block clusters of csize consecutive blocks
void launch_cluster(dim3 grid, dim3 block, size_t smem, unsigned csize, const std::function<void()>& body);
unsigned cluster_rank();
unsigned cluster_size();
void sync_cluster();
}  // namespace emu

#define threadIdx (emu::tid())
#define blockIdx (emu::bid())
#define blockDim (emu::bdim())
#define gridDim (emu::gdim())
static const int warpSize = 32;

inline void __syncthreads() { emu::sync_block(); }
inline void __syncwarp(unsigned = 0xffffffffu)
{
  uint64_t o[32];
  emu::warp_exchange(0, o);
}
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline void __threadfence_block() {}

inline unsigned __ballot_sync(unsigned, int pred)
{
  uint64_t o[32];
  emu::warp_exchange(pred ? 1 : 0, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    r |= unsigned(o[i] & 1) << i;
  return r;
}
inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
inline int __all_sync(unsigned m, int pred) { return __ballot_sync(m, pred) == 0xffffffffu; }

template <typename T>
inline uint64_t emu_pack(T v)
{
  static_assert(sizeof(T) <= 8, "");
  uint64_t u = 0;
  std::memcpy(&u, &v, sizeof(T));
  return u;
}
template <typename T>
inline T emu_unpack(uint64_t u)
{
  T v;
  std::memcpy(&v, &u, sizeof(T));
  return v;
}
template <typename T>
inline T __shfl_sync(unsigned, T v, int src, int width = 32)
{
  uint64_t o[32];
  emu::warp_exchange(emu_pack(v), o);
  int l = int(emu::lane());
  int base = l - (l % width);
  return emu_unpack<T>(o[base + (src % width)]);
}
template <typename T>
inline T __shfl_up_sync(unsigned, T v, unsigned d, int width = 32)
{
  uint64_t o[32];
  emu::warp_exchange(emu_pack(v), o);
  int l = int(emu::lane());
  int base = l - (l % width);
  int s = l - int(d);
  return s < base ? v : emu_unpack<T>(o[s]);
}
template <typename T>
inline T __shfl_down_sync(unsigned, T v, unsigned d, int width = 32)
{
  uint64_t o[32];
  emu::warp_exchange(emu_pack(v), o);
  int l = int(emu::lane());
  int base = l - (l % width);
  int s = l + int(d);
  return s >= base + width ? v : emu_unpack<T>(o[s]);
}
template <typename T>
inline T __shfl_xor_sync(unsigned, T v, int m, int width = 32)
{
  uint64_t o[32];
  emu::warp_exchange(emu_pack(v), o);
  int l = int(emu::lane());
  (void)width;
  return emu_unpack<T>(o[l ^ m]);
}
inline unsigned __match_any_sync(unsigned, unsigned v)
{
  uint64_t o[32];
  emu::warp_exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (unsigned(o[i]) == v)
      r |= 1u << i;
  return r;
}
inline unsigned __reduce_or_sync(unsigned, unsigned v)
{
  uint64_t o[32];
  emu::warp_exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    r |= unsigned(o[i]);
  return r;
}
inline unsigned __reduce_add_sync(unsigned, unsigned v)
{
  uint64_t o[32];
  emu::warp_exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    r += unsigned(o[i]);
  return r;
}
inline int __reduce_max_sync(unsigned, int v)
{
  uint64_t o[32];
  emu::warp_exchange(uint64_t(uint32_t(v)), o);
  int r = int(uint32_t(o[0]));
  for (int i = 1; i < 32; i++)
    r = std::max(r, int(uint32_t(o[i])));
  return r;
}
inline unsigned __reduce_max_sync(unsigned, unsigned v)
{
  uint64_t o[32];
  emu::warp_exchange(v, o);
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    r = std::max(r, unsigned(o[i]));
  return r;
}
inline unsigned __reduce_min_sync(unsigned, unsigned v)
{
  uint64_t o[32];
  emu::warp_exchange(v, o);
  unsigned r = 0xffffffffu;
  for (int i = 0; i < 32; i++)
    r = std::min(r, unsigned(o[i]));
  return r;
}

// bit intrinsics
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
inline int __clzll(long long v) { return v ? __builtin_clzll((unsigned long long)v) : 64; }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned __brev(unsigned v)
{
  unsigned r = 0;
  for (int i = 0; i < 32; i++)
    if (v & (1u << i))
      r |= 1u << (31 - i);
  return r;
}
inline unsigned __funnelshift_r(unsigned lo, unsigned hi, unsigned s)
{
  uint64_t v = (uint64_t(hi) << 32) | lo;
  return unsigned(v >> (s & 31));
}

// strict fp64 (this translation unit is compiled with -ffp-contract=off)
inline double __dmul_rn(double a, double b) { return a * b; }
inline double __dadd_rn(double a, double b) { return a + b; }
inline double __dsub_rn(double a, double b) { return a - b; }
inline double __ddiv_rn(double a, double b) { return a / b; }
inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
inline long long __double2ll_rn(double a) { return std::llrint(a); }
inline double __ll2double_rn(long long a) { return double(a); }
inline double __ull2double_rn(unsigned long long a) { return double(a); }
inline double __uint2double_rn(unsigned a) { return double(a); }
inline float __double2float_rn(double a) { return float(a); }
inline long long __double_as_longlong(double a) { return emu_unpack<long long>(emu_pack(a)); }
inline double __longlong_as_double(long long a) { return emu_unpack<double>(emu_pack(a)); }
template <typename T>
inline T __ldg(const T* p)
{
  return *p;
}
template <typename T>
inline T __ldcg(const T* p)
{
  if constexpr (sizeof(T) <= 8)
    return __atomic_load_n(p, __ATOMIC_RELAXED);
  else
    return *p;
}
template <typename T>
inline void __stcg(T* p, T v)
{
  __atomic_store_n(p, v, __ATOMIC_RELAXED);
}

// atomics (blocks may run on different OS threads)
template <typename T>
inline T atomicAdd(T* p, T v)
{
  return __atomic_fetch_add(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicOr(T* p, T v)
{
  return __atomic_fetch_or(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicAnd(T* p, T v)
{
  return __atomic_fetch_and(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicExch(T* p, T v)
{
  return __atomic_exchange_n(p, v, __ATOMIC_RELAXED);
}
template <typename T>
inline T atomicMax(T* p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}
template <typename T>
inline T atomicMin(T* p, T v)
{
  T old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old > v && !__atomic_compare_exchange_n(p, &old, v, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}
template <typename T>
inline T atomicCAS(T* p, T cmp, T v)
{
  __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED);
  return cmp;
}

inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline long long min(long long a, long long b) { return a < b ? a : b; }
inline long long max(long long a, long long b) { return a > b ? a : b; }
inline unsigned long long min(unsigned long long a, unsigned long long b) { return a < b ? a : b; }
inline unsigned long long max(unsigned long long a, unsigned long long b) { return a > b ? a : b; }
inline unsigned long min(unsigned long a, unsigned long b) { return a < b ? a : b; }
inline unsigned long max(unsigned long a, unsigned long b) { return a > b ? a : b; }
