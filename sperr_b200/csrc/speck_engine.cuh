// Tree-agnostic part of the integer SPECK encoder (shared by the 3D coefficient coder and the 1D
// outlier coder). See speck3d_enc.cu for the idea. A "tree policy" T supplies:
//   struct T::Data                                             device pointers describing the tree
//   static __device__ void  T::pd(data, chunk, c, node, int& p, unsigned& d)
//   static __device__ int   T::children(data, chunk, c, node, ChildRec out[8])
//     (number of children; | 0x100 when even the last child is tested explicitly)
//   T::kHasLeaf8 / T::leaf8(data, chunk, node, int p[8], unsigned& signs): optional shortcut for a set
//     whose children are all single coefficients (their msb positions and sign bits, coding order)
// where p = msb position of the largest magnitude below the node (-1: all zero) and d = number of
// bits the node's depth-first expansion emits in the plane it turns significant.
#pragma once

#include "speck.h"

namespace sperr_b200 {

struct ChildRec {
  node_t id;
  int p;
  unsigned d;         // sets only
  int kind;           // 0 pixel, 1 set whose children are all pixels, 2 any other set
  unsigned lis_desc;  // sets only: (number of lists - 1 - list index)
  unsigned sign;      // pixels only: sign bit (1 = non-negative)
};

__device__ __forceinline__ void put_bit(uint32_t* words, unsigned long long pos, unsigned bit)
{
  if (bit)
    atomicOr(&words[pos >> 5], 1u << (pos & 31));
}

// Bits a thread emits at consecutive positions: gathered in a register, OR-ed into the stream one
// word at a time (a set's test / sign bits are contiguous between the bodies of its larger children).
struct BitRun {
  uint32_t* w;
  unsigned long long pos;   // position of bit 0 of acc
  unsigned acc;
  int n;
  __device__ __forceinline__ void start(uint32_t* words, unsigned long long p)
  {
    w = words; pos = p; acc = 0; n = 0;
  }
  __device__ __forceinline__ unsigned long long cur() const { return pos + n; }
  __device__ __forceinline__ void flush()
  {
    if (acc) {
      const unsigned sh = unsigned(pos & 31);
      atomicOr(&w[pos >> 5], acc << sh);
      if (sh && (acc >> (32 - sh)))
        atomicOr(&w[(pos >> 5) + 1], acc >> (32 - sh));
    }
    pos += n;
    acc = 0;
    n = 0;
  }
  __device__ __forceinline__ void push(unsigned bit)
  {
    acc |= bit << n;
    if (++n == 32)
      flush();
  }
  __device__ __forceinline__ void skip(unsigned long long d)   // bits written by someone else
  {
    flush();
    pos += d;
  }
};

// ---------------------------------------------------------------------------------------------
// 2. LIP / refinement parts: per-plane raster-order counts, scans and emission
// ---------------------------------------------------------------------------------------------

constexpr int kLrBlock = 1024;
// SPERR_EMIT_AGG=1: a warp writes its refinement bits of a plane as one or two words instead of one
// atomic per set bit. Measured on B200 (1024^3): enc.lipref_emit 8.03 ms with it, 7.37 ms without --
// the kernel is bound by its instruction count, not by the atomics, and the two warp reductions
// cost more than the atomics they save -- so it is off.
#ifndef SPERR_EMIT_AGG
#define SPERR_EMIT_AGG 0
#endif

__device__ __forceinline__ unsigned long long load_mag(const ChunkDev& ch, unsigned long long i)
{
  return ch.wide ? reinterpret_cast<const unsigned long long*>(gptr(ch.mag))[i]
                 : (unsigned long long)reinterpret_cast<const unsigned*>(gptr(ch.mag))[i];
}

// A CUDA block takes kLrUnits consecutive 1024-coefficient units: their loads are in flight together
// and units in which no coefficient was ever created as a pixel before plane 0 (the fine sub-bands
// of a smooth field) have no LIP or refinement bits and cost no barrier.
constexpr int kLrUnits = 8;

// per coefficient: msb position | creation plane << 8 (both -1 when absent)
__device__ __forceinline__ int lr_load(const ChunkDev& ch, unsigned long long i)
{
  if (i >= ch.n)
    return -1;
  return (int(gptr(ch.pleaf)[i]) & 0xff) | (int(gptr(ch.cmap)[i]) << 8);
}
__device__ __forceinline__ int lr_p(int v) { return int(int8_t(v & 0xff)); }
__device__ __forceinline__ int lr_cm(int v) { return v >> 8; }

// counts[((c*2 + part) * maxp + n) * nblk + blk], part 0 = LIP, 1 = refinement (zeroed by the caller)
static __global__ void __launch_bounds__(kLrBlock, 2) k_lipref_count(const ChunkDev* chunks, unsigned* counts, int maxp, unsigned nblk)
{
  __shared__ unsigned s_lip[64], s_ref[64];
  __shared__ int s_cmax;
  __shared__ unsigned s_mask;
  const unsigned c = blockIdx.y, blk0 = blockIdx.x * kLrUnits;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || ch.planes == 0 || (unsigned long long)blk0 * kLrBlock >= ch.n)
    return;
  int vu[kLrUnits];
#pragma unroll
  for (int u = 0; u < kLrUnits; u++)
    vu[u] = lr_load(ch, (unsigned long long)(blk0 + u) * kLrBlock + threadIdx.x);
  unsigned live = 0;
#pragma unroll
  for (int u = 0; u < kLrUnits; u++)
    live |= lr_cm(vu[u]) > 0 ? 1u << u : 0u;   // bits exist only in planes below the creation plane
  if (threadIdx.x == 0)
    s_mask = 0;
  __syncthreads();
  live = __reduce_or_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && live)
    atomicOr(&s_mask, live);
  __syncthreads();
  const unsigned mask = s_mask;
#pragma unroll
  for (int u = 0; u < kLrUnits; u++) {
    if (!((mask >> u) & 1u))
      continue;   // uniform over the block
    const unsigned blk = blk0 + u;
    const int p = lr_p(vu[u]), cm = lr_cm(vu[u]);
    if (threadIdx.x < 64) {
      s_lip[threadIdx.x] = 0;
      s_ref[threadIdx.x] = 0;
    }
    if (threadIdx.x == 0)
      s_cmax = -1;
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, cm);
    if ((threadIdx.x & 31) == 0 && wmax >= 0)
      atomicMax(&s_cmax, wmax);
    __syncthreads();
    const int cmax = s_cmax;  // bits exist only for planes n < cmax
    for (int n = 0; n < wmax; n++) {   // warp-uniform: this warp has no bits in planes >= its largest cm
      const bool inlip = cm > n && p <= n;
      const unsigned b0 = __ballot_sync(0xffffffffu, inlip);
      const unsigned b1 = __ballot_sync(0xffffffffu, inlip && p == n);
      const unsigned b2 = __ballot_sync(0xffffffffu, p > n);
      if ((threadIdx.x & 31) == 0) {
        const unsigned l = __popc(b0) + __popc(b1), r = __popc(b2);
        if (l)
          atomicAdd(&s_lip[n], l);
        if (r)
          atomicAdd(&s_ref[n], r);
      }
    }
    __syncthreads();
    if (int(threadIdx.x) < cmax) {
      const int n = threadIdx.x;
      counts[((size_t)(c * 2 + 0) * maxp + n) * nblk + blk] = s_lip[n];
      counts[((size_t)(c * 2 + 1) * maxp + n) * nblk + blk] = s_ref[n];
    }
  }
}

// One block per (chunk, part, plane) row: exclusive scan over blocks in place, row total to sizes.
static __global__ void k_lipref_scan(const ChunkDev* chunks, unsigned* counts, unsigned long long* sizes,
                              int maxp, unsigned nblk)
{
  __shared__ unsigned wsum[32];
  __shared__ unsigned carry_s;
  const unsigned c = blockIdx.y, part = blockIdx.x / maxp, n = blockIdx.x % maxp;
  const ChunkDev& ch = chunks[c];
  unsigned long long* out = &sizes[(size_t)(c * 2 + part) * maxp + n];
  if (ch.is_const || int(n) >= ch.planes) {
    if (threadIdx.x == 0)
      *out = 0;
    return;
  }
  unsigned* row = counts + ((size_t)(c * 2 + part) * maxp + n) * nblk;
  const unsigned used = unsigned((ch.n + kLrBlock - 1) / kLrBlock);
  if (threadIdx.x == 0)
    carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (unsigned base = 0; base < used; base += blockDim.x) {
    const unsigned i = base + threadIdx.x;
    const unsigned v = i < used ? row[i] : 0;
    unsigned inc = v;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o)
        inc += t;
    }
    if (lane == 31)
      wsum[warp] = inc;
    __syncthreads();
    if (warp == 0) {
      unsigned w = wsum[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, w, o);
        if (lane >= o)
          w += t;
      }
      wsum[lane] = w;
    }
    __syncthreads();
    const unsigned carry = carry_s;
    const unsigned excl = carry + inc - v + (warp ? wsum[warp - 1] : 0);
    if (i < used)
      row[i] = excl;
    __syncthreads();
    if (threadIdx.x == 0)
      carry_s = carry + wsum[31];
    __syncthreads();
  }
  if (threadIdx.x == 0)
    *out = carry_s;
}

// bases[(c*2 + part) * maxp + n]: absolute bit position of that part (set by k_plane_begin)
static __global__ void __launch_bounds__(kLrBlock, 2) k_lipref_emit(const ChunkDev* chunks, const unsigned* counts,
                              const unsigned long long* bases, int maxp, unsigned nblk)
{
  // per plane: bits every warp of the unit contributes to the LIP part / the refinement part
  __shared__ unsigned short s_lip[64][32], s_ref[64][32];
  // absolute position of the unit's first LIP / refinement bit of every plane
  __shared__ unsigned long long s_blip[64], s_bref[64];
  __shared__ int s_cmax;
  __shared__ unsigned s_mask;
  const unsigned c = blockIdx.y, blk0 = blockIdx.x * kLrUnits;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || ch.planes == 0 || (unsigned long long)blk0 * kLrBlock >= ch.n)
    return;
  int vu[kLrUnits];
#pragma unroll
  for (int u = 0; u < kLrUnits; u++)
    vu[u] = lr_load(ch, (unsigned long long)(blk0 + u) * kLrBlock + threadIdx.x);
  const int first = ch.last_plane;  // planes below this were never coded
  const bool no_ref_last = ch.stop_after_sort != 0;
  unsigned live = 0;
#pragma unroll
  for (int u = 0; u < kLrUnits; u++)
    live |= lr_cm(vu[u]) > first ? 1u << u : 0u;   // bits exist only in coded planes below the creation plane
  if (threadIdx.x == 0)
    s_mask = 0;
  __syncthreads();
  live = __reduce_or_sync(0xffffffffu, live);
  if ((threadIdx.x & 31) == 0 && live)
    atomicOr(&s_mask, live);
  __syncthreads();
  const unsigned mask = s_mask;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
#pragma unroll
  for (int u = 0; u < kLrUnits; u++) {
    if (!((mask >> u) & 1u))
      continue;   // uniform over the block
    const unsigned blk = blk0 + u;
    const unsigned long long i = (unsigned long long)blk * kLrBlock + threadIdx.x;
    const bool valid = i < ch.n;
    const int p = lr_p(vu[u]), cm = lr_cm(vu[u]);
    if (int(threadIdx.x) < 2 * 64) {   // one global round trip per unit instead of two per plane and warp
      const int part = threadIdx.x >> 6, n = threadIdx.x & 63;
      if (n < ch.planes && n < maxp) {
        const unsigned long long b = bases[(size_t)(c * 2 + part) * maxp + n] +
                                     counts[((size_t)(c * 2 + part) * maxp + n) * nblk + blk];
        if (part == 0)
          s_blip[n] = b;
        else
          s_bref[n] = b;
      }
    }
    if (threadIdx.x == 0)
      s_cmax = -1;
    __syncthreads();
    const int wmax = __reduce_max_sync(0xffffffffu, cm);   // a coefficient has bits in planes < its cm
    if (lane == 0 && wmax >= 0)
      atomicMax(&s_cmax, wmax);
    __syncthreads();
    const int cmax = s_cmax;
    for (int n = first; n < cmax; n++) {
      unsigned l = 0, r = 0;
      if (n < wmax) {
        const bool inlip = cm > n && p <= n;
        const unsigned b0 = __ballot_sync(0xffffffffu, inlip);
        const unsigned b1 = __ballot_sync(0xffffffffu, inlip && p == n);
        const unsigned b2 = __ballot_sync(0xffffffffu, p > n && !(n == first && no_ref_last));
        l = unsigned(__popc(b0) + __popc(b1));
        r = unsigned(__popc(b2));
      }
      if (lane == 0) {
        s_lip[n][warp] = (unsigned short)l;
        s_ref[n][warp] = (unsigned short)r;
      }
    }
    __syncthreads();
    const unsigned long long mag = (valid && p >= 0 && wmax >= 0) ? load_mag(ch, i) : 0;
    const unsigned sgn = (valid && wmax >= 0) ? (gptr(ch.signs)[i >> 5] >> (i & 31)) & 1u : 0;
    for (int n = wmax - 1; n >= first; n--) {   // warp-uniform bounds (no iteration when wmax < 0)
      const bool inlip = cm > n && p <= n;
      const bool newsig = inlip && p == n;
      const bool ref = p > n && !(n == first && no_ref_last);
      const unsigned b0 = __ballot_sync(0xffffffffu, inlip);
      const unsigned b1 = __ballot_sync(0xffffffffu, newsig);
      const unsigned b2 = __ballot_sync(0xffffffffu, ref);
      if (b0) {
        const unsigned before = __reduce_add_sync(0xffffffffu, lane < warp ? unsigned(s_lip[n][lane]) : 0u);
        if (inlip) {
          const unsigned long long pos = s_blip[n] + before + __popc(b0 & lt) + __popc(b1 & lt);
          if (newsig) {
            put_bit(gptr(ch.spk), pos, 1);
            put_bit(gptr(ch.spk), pos + 1, sgn);
          }
        }
      }
#if SPERR_EMIT_AGG
      if (b2) {
        // the warp's refinement bits of this plane are one contiguous run of at most 32 bits: OR-ed
        // together across the warp and written as one or two words (one atomic per word instead of
        // one per set bit -- the staging words are shared with the warps around this one)
        const unsigned before = __reduce_add_sync(0xffffffffu, lane < warp ? unsigned(s_ref[n][lane]) : 0u);
        const unsigned long long start = s_bref[n] + before;   // warp-uniform
        const unsigned off = unsigned(start & 31) + __popc(b2 & lt);   // < 63
        const unsigned bit = ref ? unsigned(mag >> n) & 1u : 0u;
        const unsigned lo = __reduce_or_sync(0xffffffffu, off < 32 ? bit << off : 0u);
        const unsigned hi = __reduce_or_sync(0xffffffffu, off >= 32 ? bit << (off - 32) : 0u);
        if (lane == 0) {
          uint32_t* const w = gptr(ch.spk) + (start >> 5);
          if (lo)
            atomicOr(w, lo);
          if (hi)
            atomicOr(w + 1, hi);
        }
      }
#else
      if (b2) {
        const unsigned before = __reduce_add_sync(0xffffffffu, lane < warp ? unsigned(s_ref[n][lane]) : 0u);
        if (ref) {
          const unsigned long long pos = s_bref[n] + before + __popc(b2 & lt);
          put_bit(gptr(ch.spk), pos, unsigned(mag >> n) & 1u);
        }
      }
#endif
    }
    // the shared arrays are rewritten by the next unit only after its first barrier ... except
    // s_blip / s_bref, which the slowest warp may still be reading:
    __syncthreads();
  }
}


// ---------------------------------------------------------------------------------------------
// LIS part, one plane at a time
// ---------------------------------------------------------------------------------------------

// step s of the bit-plane loop: which plane does each chunk code?
static __global__ void k_plane_pre(EncCtx ctx, int step)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ctx.nchunks)
    return;
  ChunkDev& ch = ctx.chunks[c];
  ch.cur_n = ch.planes - 1 - step;
  ch.plane_live = ch.active && ch.cur_n >= 0;
  ch.ncand = 0;
}

template <class T>
__global__ void k_root_seglen(EncCtx ctx, typename T::Data tree, unsigned long long nroots)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nroots;
       i += stride) {
    const unsigned c = key_chunk(ctx.rkey[i]);
    const ChunkDev& ch = ctx.chunks[c];
    int p;
    unsigned d;
    T::pd(tree, ch, c, ctx.rnode[i], p, d);
    ctx.rseg[i] = 1 + (p == ch.cur_n ? d : 0u);
  }
}

// Places the three parts of this plane and applies the fixed-rate budget rules
// (src/SPECK_INT.cpp:145-158).
static __global__ void k_plane_begin(EncCtx ctx)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ctx.nchunks)
    return;
  ChunkDev& ch = ctx.chunks[c];
  ch.next_needed = 0;
  if (!ch.plane_live)
    return;
  const int n = ch.cur_n;
  const unsigned long long lis = ctx.rpos[ctx.rstart[c + 1]] - ctx.rpos[ctx.rstart[c]];
  const unsigned long long lip = ctx.sizes[(size_t)(c * 2 + 0) * ctx.maxp + n];
  const unsigned long long ref = ctx.sizes[(size_t)(c * 2 + 1) * ctx.maxp + n];
  ctx.bases[(size_t)(c * 2 + 0) * ctx.maxp + n] = ch.cursor;
  ch.lis_base = ch.cursor + lip;
  const unsigned long long after_sort = ch.lis_base + lis;
  ctx.bases[(size_t)(c * 2 + 1) * ctx.maxp + n] = after_sort;
  ch.last_plane = n;
  if (after_sort >= ch.budget) {
    ch.stop_after_sort = 1;
    ch.total_bits = after_sort;
    ch.active = 0;
    return;
  }
  const unsigned long long after_ref = after_sort + ref;
  ch.total_bits = after_ref;
  if (after_ref >= ch.budget || n == 0) {
    ch.active = 0;
    return;
  }
  ch.cursor = after_ref;
  ch.next_needed = 1;
}

// ---- work-list appends: one atomic per warp ---------------------------------------------------
// Every lane of the warp calls with the number of slots it needs; returns the lane's first slot.
__device__ __forceinline__ unsigned long long warp_reserve(unsigned long long* counter, unsigned n)
{
  const int lane = threadIdx.x & 31;
  unsigned inc = n;
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o)
      inc += t;
  }
  const unsigned total = __shfl_sync(0xffffffffu, inc, 31);
  unsigned long long base = 0;
  if (lane == 31 && total)
    base = atomicAdd(counter, (unsigned long long)total);
  base = __shfl_sync(0xffffffffu, base, 31);
  return base + inc - n;
}

// ---- grid-wide barrier of the persistent plane kernel (all CTAs are co-resident: cooperative
// launch). bar[0]: arrivals, bar[1]: generation. ----
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned nblocks)
{
  cta_sync();
  if (threadIdx.x == 0 && nblocks > 1) {
    __threadfence();
    volatile unsigned* const vbar = bar;   // volatile: re-read from L2 on every turn of the spin
    const unsigned gen = vbar[1];
    if (atomicAdd(&bar[0], 1u) == nblocks - 1) {
      atomicExch(&bar[0], 0u);
      __threadfence();
      atomicAdd(&bar[1], 1u);
    }
    else {
      unsigned spins = 0;
      while (vbar[1] == gen) {
#if !defined(SPERR_EMUL)
        __nanosleep(64);
        if (++spins > (1u << 26))   // seconds: a CTA is missing, fail the launch instead of hanging
          __trap();
#endif
      }
      (void)spins;
    }
    __threadfence();
  }
  cta_sync();   // thread 0 arrives from its spin loop: not the aligned form
}

// The LIS part of one plane in ONE launch: the roots' significance bits, then the expansion of the
// significant sets level by level until the frontier is empty -- the number of rounds is whatever
// the data needs, decided on the device. m_code_S (src/SPECK3D_INT.cpp:140-212,
// src/SPECK1D_INT_ENC.cpp:121-160) with every child placed by its D instead of by recursion order.
// Frontier buffers alternate (round & 1); their counters rotate over three slots so that the one a
// round appends to was cleared a full round earlier.
template <class T>
__global__ void __launch_bounds__(256) k_lis_plane(EncCtx ctx, typename T::Data tree, unsigned* bar)
{
  const unsigned nblocks = gridDim.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const unsigned long long tid0 = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned long long nroots = *ctx.total_roots;
  // --- roots ---
  for (unsigned long long b0 = 0; b0 < nroots; b0 += stride) {   // warp-uniform trip count
    const unsigned long long i = b0 + tid0;
    bool sig = false, keep = false;
    unsigned long long key = 0, pos = 0;
    node_t nd = 0;
    unsigned c = 0;
    if (i < nroots) {
      key = ctx.rkey[i];
      c = key_chunk(key);
      const ChunkDev& ch = ctx.chunks[c];
      if (ch.plane_live) {
        pos = ch.lis_base + (ctx.rpos[i] - ctx.rpos[ctx.rstart[c]]);
        nd = ctx.rnode[i];
        int p;
        unsigned d;
        T::pd(tree, ch, c, nd, p, d);
        if (p == ch.cur_n) {
          put_bit(gptr(ch.spk), pos, 1);
          sig = true;
        }
        else
          keep = ch.next_needed != 0;   // insignificant: stays in its list, same position
      }
    }
    const unsigned long long fs = warp_reserve(&ctx.fcount[0], sig ? 1u : 0u);
    const unsigned long long cs = warp_reserve(ctx.cand_count, keep ? 1u : 0u);
    if (sig) {
      if (fs < ctx.front_cap) {
        ctx.fnode[0][fs] = nd;
        ctx.fpos[0][fs] = ((pos + 1) << 10) | c;
      }
      else
        atomicOr(ctx.err, 2u);
    }
    if (keep) {
      if (cs < ctx.cand_cap) {
        ctx.ckey[cs] = key;
        ctx.cnode[cs] = nd;
      }
      else
        atomicOr(ctx.err, 1u);
    }
  }
  grid_barrier(bar, nblocks);
  // --- expansion rounds ---
  for (int round = 0;; round++) {
    const int src = round & 1, dst = src ^ 1;
    unsigned long long* const cnt_src = &ctx.fcount[round % 3];
    unsigned long long* const cnt_dst = &ctx.fcount[(round + 1) % 3];
    const unsigned long long count = min(__ldcg(cnt_src), ctx.front_cap);
    if (count == 0)
      break;   // the same on every thread of the grid
    if (tid0 == 0)
      ctx.fcount[(round + 2) % 3] = 0;   // read a round ago, appended to a round from now
    for (unsigned long long b0 = 0; b0 < count; b0 += stride) {
      const unsigned long long i = b0 + tid0;
      const bool valid = i < count;
      node_t nd = 0;
      unsigned c = 0;
      unsigned long long cur = 0;
      ChildRec kid[8];
      int nch = 0, n = 0;
      bool all_tested = false, next_needed = false;
      unsigned ncand = 0, nfront = 0;
      if (valid) {
        nd = ctx.fnode[src][i];
        const unsigned long long pc = ctx.fpos[src][i];
        c = unsigned(pc & 1023u);
        cur = pc >> 10;
        const ChunkDev& ch = ctx.chunks[c];
        n = ch.cur_n;
        next_needed = ch.next_needed != 0;
        const int nchf = T::children(tree, ch, c, nd, kid);
        nch = nchf & 0xff;
        all_tested = (nchf & 0x100) != 0;   // 2D: last split of the set I
        int sc = 0;
        for (int k = 0; k < nch; k++) {   // what will this set append?
          const bool need = sc != 0 || k != nch - 1 || all_tested;
          const bool sig = !need || kid[k].p == n;
          sc += sig;
          if (kid[k].kind != 0) {
            if (sig)
              nfront += kid[k].kind == 2;
            else
              ncand += next_needed;
          }
        }
      }
      unsigned long long fs = warp_reserve(cnt_dst, nfront);
      unsigned long long cs = warp_reserve(ctx.cand_count, ncand);
      if (!valid)
        continue;
      if (fs + nfront > ctx.front_cap) {
        atomicOr(ctx.err, 2u);
        continue;
      }
      if (cs + ncand > ctx.cand_cap) {
        atomicOr(ctx.err, 1u);
        continue;
      }
      const ChunkDev& ch = ctx.chunks[c];
      BitRun run;
      run.start(gptr(ch.spk), cur);
      int sigc = 0;
      for (int k = 0; k < nch; k++) {
        const bool need = sigc != 0 || k != nch - 1 || all_tested;
        const bool sig = !need || kid[k].p == n;
        if (need)
          run.push(sig);
        if (kid[k].kind == 0) {
          if (sig) {
            run.push(kid[k].sign);
            sigc++;
          }
        }
        else if (sig) {
          sigc++;
          if (kid[k].kind == 1) {  // all grandchildren are pixels: finish them here
            int gp[8];
            unsigned gsign = 0;
            int ng;
            if constexpr (T::kHasLeaf8)
              ng = T::leaf8(tree, ch, kid[k].id, gp, gsign);
            else {
              ChildRec g[8];
              ng = T::children(tree, ch, c, kid[k].id, g) & 0xff;
              for (int j = 0; j < ng; j++) {
                gp[j] = g[j].p;
                gsign |= g[j].sign << j;
              }
            }
            int gs = 0;
            for (int j = 0; j < ng; j++) {
              const bool gneed = gs != 0 || j != ng - 1;
              const bool gsig = !gneed || gp[j] == n;
              if (gneed)
                run.push(gsig);
              if (gsig) {
                run.push((gsign >> j) & 1u);
                gs++;
              }
            }
          }
          else {
            ctx.fnode[dst][fs] = kid[k].id;
            ctx.fpos[dst][fs] = (run.cur() << 10) | c;
            fs++;
            run.skip(kid[k].d);
          }
        }
        else if (next_needed) {
          ctx.ckey[cs] = make_key(c, kid[k].lis_desc, (run.cur() - 1) + 64);
          ctx.cnode[cs] = kid[k].id;
          cs++;
        }
      }
      run.flush();
    }
    grid_barrier(bar, nblocks);
  }
}

// After the live sets have been sorted into list order: first root of every chunk, counters cleared.
static __global__ void k_plane_end(EncCtx ctx, unsigned long long total)
{
  for (int c = threadIdx.x; c <= ctx.nchunks; c += blockDim.x) {
    // first key whose chunk is >= c
    unsigned long long lo = 0, hi = total;
    while (lo < hi) {
      const unsigned long long mid = (lo + hi) >> 1;
      if (key_chunk(ctx.rkey[mid]) < unsigned(c))
        lo = mid + 1;
      else
        hi = mid;
    }
    ctx.rstart[c] = c == ctx.nchunks ? total : lo;
  }
  if (threadIdx.x == 0) {
    *ctx.total_roots = total;
    *ctx.cand_count = 0;
    ctx.fcount[0] = 0;
    ctx.fcount[1] = 0;
    ctx.fcount[2] = 0;
  }
}

// Seeds the per-chunk coder state once `planes` is known and lays the initial sets into the
// sorted root arrays. `nroots[c]` initial sets per chunk are produced by T::root(data, c, r, ...).
template <class T>
__global__ void k_enc_seed(EncCtx ctx, typename T::Data tree)
{
  __shared__ unsigned long long s_start[kMaxBatchChunks + 1];
  for (int c = threadIdx.x; c < ctx.nchunks; c += blockDim.x) {
    ChunkDev& ch = ctx.chunks[c];
    const int P = ch.is_const ? 0 : T::planes(tree, ch, c);
    ch.planes = P;
    ch.active = P > 0;
    ch.cursor = 0;
    ch.total_bits = 0;
    ch.last_plane = 0;
    ch.stop_after_sort = 0;
    ch.ncand = 0;
    s_start[c] = P > 0 ? T::num_roots(tree, ch, c) : 0;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int c = 0; c < ctx.nchunks; c++) {
      const unsigned long long v = s_start[c];
      s_start[c] = acc;
      acc += v;
    }
    s_start[ctx.nchunks] = acc;
    *ctx.total_roots = acc;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= ctx.nchunks; c += blockDim.x)
    ctx.rstart[c] = s_start[c];
  for (int c = threadIdx.x; c < ctx.nchunks; c += blockDim.x) {
    const ChunkDev& ch = ctx.chunks[c];
    if (ch.planes == 0)
      continue;
    const int nr = T::num_roots(tree, ch, c);
    for (int r = 0; r < nr; r++) {
      unsigned lis_desc, order;
      node_t nd;
      T::root(tree, ch, c, r, nd, lis_desc, order);
      ctx.rkey[s_start[c] + r] = make_key(c, lis_desc, order);
      ctx.rnode[s_start[c] + r] = nd;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// host driver of the bit-plane loop
// ---------------------------------------------------------------------------------------------

// One cooperative launch (every CTA resident, so the grid barrier cannot deadlock): as many CTAs of
// 256 threads as the device holds at once, at most 8 per SM.
template <class T>
void launch_lis_plane(const EncCtx& ctx, const typename T::Data& tree, unsigned* d_bar, cudaStream_t st)
{
#ifdef SPERR_EMUL
  LAUNCH(k_lis_plane<T>, dim3(1), dim3(256), 0, st, ctx, tree, d_bar);
#else
  static int grids[rt::kMaxDevices] = {};
  int& grid = grids[rt::cur_dev()];
  if (!grid) {
    int dev = 0, sms = 0, per = 0;
    RT_CHECK(cudaGetDevice(&dev));
    RT_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    RT_CHECK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, k_lis_plane<T>, 256, 0));
    // The compressor runs two of these loops side by side (SPECK3D of the coefficients, SPECK1D of the
    // outliers), and a cooperative grid only starts when ALL its CTAs fit: sized to the whole GPU
    // each, the two chains take turns (measured: c.outlier_encode 6 -> 25 ms). The big tree takes
    // all but one CTA slot per SM, the outlier tree one.
    const int slots = std::max(1, std::min(per, 8));
    grid = T::kIsOutlierTree ? sms : sms * std::max(1, slots - 1);
  }
  EncCtx c = ctx;
  typename T::Data t = tree;
  void* args[] = {&c, &t, &d_bar};
  rt::check(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(&k_lis_plane<T>), dim3(unsigned(grid)),
                                        dim3(256), args, 0, st),
            "k_lis_plane", __FILE__, __LINE__);
  rt::launch_counter()++;
#endif
}

// `cap_nodes`: upper bound on the number of sets alive / expanded at once over the whole batch.
// `bits_bound(c, planes)`: upper bound of the payload size of chunk c. `max_depth`: number of
// expansion rounds that empties any frontier.
template <class T, class BitsBound>
void run_encoder(EncWork& w, ChunkDev* d_chunks, int nchunks, size_t max_n,
                 const typename T::Data& tree, unsigned long long cap_nodes, int max_depth,
                 BitsBound bits_bound, std::vector<EncResult>& results, cudaStream_t st)
{
  if (nchunks > kMaxBatchChunks)
    throw std::runtime_error("too many chunks in one batch");
  results.assign(nchunks, EncResult());
  EncCtx ctx;
  ctx.chunks = d_chunks;
  ctx.nchunks = nchunks;
  ctx.shapes = nullptr;
  ctx.cand_cap = cap_nodes + 64ull * nchunks;
  ctx.front_cap = ctx.cand_cap;
  const size_t cap = size_t(ctx.cand_cap);
  for (int i = 0; i < 2; i++) {
    w.keys[i].reserve(cap * 8);
    w.nodes[i].reserve(cap * 8);
    w.fnode[i].reserve(cap * 8);
    w.fpos[i].reserve(cap * 8);
  }
  w.rseg.reserve((cap + 1) * 4);
  w.rpos.reserve((cap + 2) * 8);
  w.scan_tmp.reserve(scan_tmp_bytes(cap + 1));
  const size_t sort_bytes = sort_tmp_bytes(cap);
  w.sort_tmp.reserve(sort_bytes);
  w.small.reserve(4096 + (kMaxBatchChunks + 1) * 8);
  rt::dset(w.small.p, 0, w.small.bytes, st);
  unsigned long long* small = w.small.as<unsigned long long>();
  ctx.total_roots = small + 0;
  ctx.cand_count = small + 1;
  ctx.fcount = small + 5;  // three rotating counters
  ctx.err = reinterpret_cast<unsigned*>(small + 4);
  unsigned* const d_bar = reinterpret_cast<unsigned*>(small + 2);   // grid barrier of k_lis_plane
  ctx.rstart = small + 8;
  ctx.rkey = w.keys[0].as<unsigned long long>();
  ctx.rnode = w.nodes[0].as<node_t>();
  ctx.ckey = w.keys[1].as<unsigned long long>();
  ctx.cnode = w.nodes[1].as<node_t>();
  ctx.fnode[0] = w.fnode[0].as<node_t>(); ctx.fnode[1] = w.fnode[1].as<node_t>();
  ctx.fpos[0] = w.fpos[0].as<unsigned long long>(); ctx.fpos[1] = w.fpos[1].as<unsigned long long>();
  ctx.rseg = w.rseg.as<unsigned>();
  ctx.rpos = w.rpos.as<unsigned long long>();
  ctx.maxp = 1;
  ctx.sizes = ctx.bases = nullptr;

  rt::ProfScope* ps_seed = new rt::ProfScope(T::kIsOutlierTree ? "enc1d.seed" : "enc.seed", st);
  LAUNCH(k_enc_seed<T>, dim3(1), dim3(1024), 0, st, ctx, tree);

  // planes are only known on the device: fetch them to size the staging buffers and tables
  std::vector<ChunkDev> hc(nchunks);
  rt::d2h(hc.data(), d_chunks, sizeof(ChunkDev) * nchunks, st);
  unsigned long long total_roots = 0;
  rt::d2h(&total_roots, ctx.total_roots, 8, st);
  rt::sync(st);
  delete ps_seed;
  rt::ProfScope* ps_zero = new rt::ProfScope(T::kIsOutlierTree ? "enc1d.stage_zero" : "enc.stage_zero", st);
  int maxp = 1;
  for (auto& c : hc)
    maxp = std::max(maxp, c.planes);
  ctx.maxp = maxp;

  // staging bit arrays
  std::vector<unsigned long long> stage_words_off;
  {
    std::vector<unsigned long long> words_off(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
      const unsigned long long bits = hc[c].planes > 0 ? bits_bound(c, hc[c]) : 0;
      words_off[c + 1] = words_off[c] + (bits + 31) / 32 + 2;
    }
    const size_t stage_bytes = words_off[nchunks] * 4;
    if (stage_bytes > w.stage.bytes)
      w.stage_clean = false;   // reserve() will hand out new, uninitialised memory
    w.stage.reserve(stage_bytes);
    if (!w.stage_clean)
      rt::dset(w.stage.p, 0, w.stage.bytes, st);
    else
      for (const auto& r : w.stage_dirty)
        rt::dset(w.stage.as<uint32_t>() + r.first, 0, r.second * 4, st);
    if (std::getenv("SPERR_B200_STAGE_DEBUG")) {
      size_t dirty = 0;
      for (const auto& r : w.stage_dirty)
        dirty += r.second * 4;
      std::fprintf(stderr, "stage: need %zu bytes, buffer %zu, %s, dirty ranges %zu (%zu bytes)\n", stage_bytes, w.stage.bytes,
                   w.stage_clean ? "clean" : "zeroed as a whole", w.stage_dirty.size(), dirty);
    }
    w.stage_clean = false;   // until this run has recorded what it wrote (an exception leaves it false)
    w.stage_dirty.clear();
    stage_words_off = words_off;
    for (int c = 0; c < nchunks; c++) {
      hc[c].spk = w.stage.as<uint32_t>() + words_off[c];
      hc[c].spk_cap_bits = (words_off[c + 1] - words_off[c]) * 32;
    }
    rt::h2d(d_chunks, hc.data(), sizeof(ChunkDev) * nchunks, st);
  }

  // LIP / refinement counts for every plane
  const unsigned nblk = unsigned((max_n + kLrBlock - 1) / kLrBlock);
  const size_t ncounts = (size_t)nchunks * 2 * maxp * std::max(nblk, 1u);
  w.counts.reserve(ncounts * 4);
  rt::dset(w.counts.p, 0, ncounts * 4, st);
  w.sizes.reserve((size_t)nchunks * 2 * maxp * 8 * 2);
  rt::dset(w.sizes.p, 0, (size_t)nchunks * 2 * maxp * 8 * 2, st);
  ctx.sizes = w.sizes.as<unsigned long long>();
  ctx.bases = ctx.sizes + (size_t)nchunks * 2 * maxp;
  unsigned* d_counts = w.counts.as<unsigned>();
  delete ps_zero;
  if (w.after_setup)
    w.after_setup(st);
  if (nblk) {
    rt::ProfScope ps(T::kIsOutlierTree ? "enc1d.lipref_count" : "enc.lipref_count", st);
    LAUNCH(k_lipref_count, dim3((nblk + kLrUnits - 1) / kLrUnits, nchunks), dim3(kLrBlock), 0, st, d_chunks, d_counts, maxp, nblk);
    LAUNCH(k_lipref_scan, dim3(2 * maxp, nchunks), dim3(1024), 0, st, d_chunks, d_counts, ctx.sizes,
           maxp, nblk);
  }

  if (w.before_plane_loop)
    w.before_plane_loop(st);

  // bit-plane loop (LIS part)
  void* d_scan_tmp = w.scan_tmp.p;
  const unsigned cgrid = unsigned((nchunks + 127) / 128);
  for (int step = 0; step < maxp; step++) {
    rt::ProfScope ps(T::kIsOutlierTree ? "enc1d.plane_loop" : "enc.plane_loop", st);
    LAUNCH(k_plane_pre, dim3(cgrid), dim3(128), 0, st, ctx, step);
    const unsigned rgrid = unsigned(std::min<unsigned long long>((total_roots + 255) / 256, 148 * 16));
    if (total_roots)
      LAUNCH(k_root_seglen<T>, dim3(rgrid), dim3(256), 0, st, ctx, tree, total_roots);
    exclusive_scan_u32(ctx.rseg, ctx.rpos, total_roots, d_scan_tmp, st);
    LAUNCH(k_plane_begin, dim3(cgrid), dim3(128), 0, st, ctx);
    if (total_roots)
      launch_lis_plane<T>(ctx, tree, d_bar, st);
    rt::d2h(&total_roots, ctx.cand_count, 8, st);
    rt::sync(st);
    if (total_roots > ctx.cand_cap)
      throw std::runtime_error("SPECK list overflow");
    // later planes may have no LIS part; k_plane_begin still places LIP / refinement
    if (total_roots)
      sort_pairs_u64(ctx.ckey, ctx.rkey, ctx.cnode, ctx.rnode, total_roots, kKeyBits, w.sort_tmp.p,
                     sort_bytes, st);
    LAUNCH(k_plane_end, dim3(1), dim3(1024), 0, st, ctx, total_roots);
  }
  (void)max_depth;

  // LIP / refinement emission
  rt::ProfScope ps_emit(T::kIsOutlierTree ? "enc1d.lipref_emit" : "enc.lipref_emit", st);
  if (nblk)
    LAUNCH(k_lipref_emit, dim3((nblk + kLrUnits - 1) / kLrUnits, nchunks), dim3(kLrBlock), 0, st, d_chunks, d_counts, ctx.bases,
           maxp, nblk);

  rt::d2h(hc.data(), d_chunks, sizeof(ChunkDev) * nchunks, st);
  unsigned err = 0;
  rt::d2h(&err, ctx.err, 4, st);
  rt::sync(st);
  if (err)
    throw std::runtime_error("SPECK encoder work-list overflow");
  // what this run wrote into the staging array: bits [0, total_bits) of every chunk's slot
  for (int c = 0; c < nchunks; c++) {
    const size_t slot = size_t(stage_words_off[c + 1] - stage_words_off[c]);
    const size_t used = hc[c].planes > 0 ? std::min<size_t>(slot, size_t((hc[c].total_bits + 31) / 32 + 2)) : 0;
    if (used)
      w.stage_dirty.emplace_back(size_t(stage_words_off[c]), used);
  }
  w.stage_clean = true;
  for (int c = 0; c < nchunks; c++) {
    results[c].planes = hc[c].planes;
    results[c].total_bits = hc[c].total_bits;
    results[c].payload = hc[c].spk;
    const unsigned long long pack = std::min(hc[c].total_bits, hc[c].budget);
    results[c].payload_bytes = size_t((pack + 7) / 8);
  }
}

}  // namespace sperr_b200
