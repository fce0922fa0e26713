// Tree-agnostic integer SPECK decoder: one CTA per chunk walks the bit-planes of its stream.
//
// Reproduces SPECK_INT<T>::decode (/root/reference/src/SPECK_INT.cpp:165-228), the decoder's
// refinement pass (:359-469) and the sorting passes of the 3D / 1D coders
// (src/SPECK3D_INT.cpp:99-212 with src/SPECK3D_INT_DEC.cpp:8-49; src/SPECK1D_INT_DEC.cpp:12-125).
//
// How the work is split inside a CTA (not how the reference does it):
//   * LIP part of a plane: the bit string is a 2-state grammar (significance bit, then a sign bit
//     iff it was 1), so it is tokenised by a block-wide prefix scan over state-transition functions;
//     results are matched to pixels by a popcount scan of the LIP mask in raster order.
//   * LIS part: the only inherently serial piece -- every bit decides what the next bit means. One
//     thread walks the lists depth-first. Lists hold only live (insignificant) sets and are
//     compacted in place while they are iterated, which keeps SPECK's list order:
//     survivors first, then the sets created in this plane in depth-first creation order.
//   * refinement part: bit k belongs to the k-th pixel of the LSP mask in raster order (popcount
//     scan); truncated streams stop exactly where the reference stops.
// A "tree policy" T supplies the initial sets and the children of a set:
//   static __device__ int  T::roots(data, dc, c, r, node_t& nd, int& lis)    r-th initial set
//   static __device__ int  T::num_roots(data, dc, c)
//   static __device__ int  T::children(data, dc, c, node, lis, DChild out[8])
//     (number of children; | 0x100 when even the last child is tested explicitly)
//   T::kHasI / T::iset(data, dc, c): 2D only, the initial set I (0: none); a child with lis < 0 is I
#pragma once

#include "speck_dec.h"
#include "speck.h"

namespace sperr_b200 {

// The serial parts of the decoders ("one thread walks ...") are run by ALL 32 lanes of warp 0,
// redundantly: the lanes hold the same values, take the same branches and store the same data to the
// same addresses, so the result is that of one thread -- but the warp never diverges. With
// `if (tid == 0) { long serial part }` the warp of thread 0 was not always reconverged at the barrier
// that follows (B200, nvcc 12.9 -O3; compute-sanitizer synccheck: "divergent thread(s) in warp"), and
// a barrier reached that way releases early: decoder state one barrier out of step. Nothing inside a
// serial part may therefore be an atomic or depend on the lane. (The CPU emulator runs lanes one
// after the other, not in lockstep: there it is thread 0 alone.)
// Lockstep is what makes a read-modify-write inside a serial part (S.klip -= t) happen once: the
// lanes are brought together when they enter (DEC_SERIAL_CONVERGE, the first statement of every
// serial part) -- per-lane branches before it, or a barrier that does not converge the warp, may have
// left them apart.
#ifdef SPERR_EMUL
#define DEC_SERIAL(tid) ((tid) == 0)
#define DEC_SERIAL_CONVERGE() ((void)0)
#else
#define DEC_SERIAL(tid) (((tid) >> 5) == 0)
#define DEC_SERIAL_CONVERGE() __syncwarp()
#endif

// CTA barrier of the decoders. SPERR_DEC_UNALIGNED_BARRIER: barrier.sync without .aligned (threads of
// a warp may arrive separately; about 30 % slower on this kernel), kept as a switch.
__device__ __forceinline__ void block_sync()
{
#ifdef SPERR_DEC_UNALIGNED_BARRIER
  cta_sync();
#else
  __syncthreads();
#endif
}

constexpr int kDecThreads = 1024;
constexpr int kDecWarps = kDecThreads / 32;

struct DecShared {
  unsigned long long pos;        // read position (bits)
  unsigned long long klip, klsp; // population of the LIP / number of significant coefficients
  unsigned long long knew;       // coefficients found significant in the current plane
  unsigned long long endpos;
  unsigned long long wtot[kDecWarps];
  unsigned long long wtot2[kDecWarps];
  unsigned fA[kDecWarps], fB[kDecWarps];
  unsigned tile_fA, tile_fB;
};

// ---- bit access -----------------------------------------------------------------------------

// 64 bits starting at bit position p; bits at or beyond `avail` read as zero.
__device__ __forceinline__ unsigned long long fetch64(const DecChunk& d, unsigned long long p)
{
  if (p >= d.avail)
    return 0ull;
  const unsigned long long w = p >> 5;
  const unsigned sh = unsigned(p & 31);
  const unsigned long long lo = gptr(d.bits)[w], mid = gptr(d.bits)[w + 1], hi = gptr(d.bits)[w + 2];
  unsigned long long v = (lo | (mid << 32)) >> sh;
  if (sh)
    v |= hi << (64 - sh);
  const unsigned long long left = d.avail - p;
  if (left < 64)
    v &= (1ull << left) - 1ull;
  return v;
}

struct BitReader {
  const uint32_t* w;
  unsigned long long avail, pos;
  unsigned long long cur_idx;
  uint32_t cur;
  __device__ __forceinline__ void init(const DecChunk& d, unsigned long long p)
  {
    w = gptr(d.bits);
    avail = d.avail;
    pos = p;
    cur_idx = ~0ull;
    cur = 0;
  }
  __device__ __forceinline__ unsigned get()
  {
    unsigned b = 0;
    if (pos < avail) {
      const unsigned long long i = pos >> 5;
      if (i != cur_idx) {
        cur_idx = i;
        cur = w[i];
      }
      b = (cur >> (pos & 31)) & 1u;
    }
    pos++;
    return b;
  }
};

// ---- transition functions of the LIP grammar ---------------------------------------------------
// state 0: next bit is a significance bit; state 1: next bit is a sign bit.
// A function is stored per entry state as (exit state | tokens << 1).

__device__ __forceinline__ unsigned lip_sim(unsigned w, int s)
{
  unsigned c = 0;
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const unsigned b = (w >> i) & 1u;
    if (s == 0) {
      c++;
      s = int(b);
    }
    else
      s = 0;
  }
  return unsigned(s) | (c << 1);
}

// (g after f)
__device__ __forceinline__ void lip_compose(unsigned fA, unsigned fB, unsigned gA, unsigned gB,
                                            unsigned& oA, unsigned& oB)
{
  const unsigned ga = (fA & 1u) ? gB : gA;
  const unsigned gb = (fB & 1u) ? gB : gA;
  oA = (ga & 1u) | (((fA >> 1) + (ga >> 1)) << 1);
  oB = (gb & 1u) | (((fB >> 1) + (gb >> 1)) << 1);
}

__device__ __forceinline__ unsigned long long warp_incl_scan(unsigned long long v, int lane)
{
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned long long t = __shfl_up_sync(0xffffffffu, v, o);
    if (lane >= o)
      v += t;
  }
  return v;
}

// ---- LIP part ---------------------------------------------------------------------------------

__device__ __forceinline__ uint4 lip_load(const DecChunk& d, unsigned long long j)
{
  const uint4* const p = reinterpret_cast<const uint4*>(gptr(d.lip) + j);
  return d.R > 1 ? __ldcg(p) : *p;
}
__device__ __forceinline__ uint32_t lip_res_load(const DecChunk& d, const uint32_t* p)
{
  return d.R > 1 ? __ldcg(p) : *p;
}
// SPERR_LIP_SUMMARY=1: the LIP pass skips the 128-word groups of the mask no pixel was ever put in
// (DecChunk::lipsum). Measured on B200 (bench field): the LIP phase 9.6 -> 8.8 Mcycles per chunk
// (11.7 -> 9.4 at 64 chunks), the expansion phase, which has to mark the groups, 15.3 -> 16.5; decode
// stage 34.0 -> 33.4 ms at 64 chunks, 19.2 -> 19.6 ms at one. The mask sweeps are not what the LIP
// phase spends its time on (the token scan over the stream is); off.
#ifndef SPERR_LIP_SUMMARY
#define SPERR_LIP_SUMMARY 0
#endif
// word `w` of the LIP mask has received a pixel
__device__ __forceinline__ void lip_mark_group(const DecChunk& d, unsigned long long w)
{
#if SPERR_LIP_SUMMARY
  const unsigned long long g = w >> 7;
  uint32_t* const p = gptr(d.lipsum) + (g >> 5);
  const uint32_t bit = 1u << (g & 31);
  if (!(__ldcg(p) & bit))   // (L2: the bit is set with an atomic, by any CTA of the cluster)
    atomicOr(p, bit);
#endif
}
// The summary bits of the groups a warp sweeps, 32 per lane (lane l: groups g0 + 32 l ...), fetched
// once per LIP pass; group number `it` of the warp's range is live iff its bit is set.
__device__ __forceinline__ uint32_t lip_summary_of(const DecChunk& d, unsigned long long g0, int lane)
{
#if SPERR_LIP_SUMMARY
  const unsigned long long g = g0 + 32ull * lane;
  const uint32_t* const p = gptr(d.lipsum) + (g >> 5);
  const unsigned long long last = (((d.n + 31) / 32) >> 7) >> 5;   // last word that holds a group
  const uint32_t lo = (g >> 5) <= last ? __ldcg(p) : 0u, hi = (g >> 5) + 1 <= last ? __ldcg(p + 1) : 0u;
  return __funnelshift_r(lo, hi, unsigned(g & 31));
#else
  return 0xffffffffu;
#endif
}
__device__ __forceinline__ bool lip_group_live(const DecChunk& d, uint32_t mine, unsigned long long g0,
                                               unsigned long long it)
{
#if SPERR_LIP_SUMMARY
  if (it < 1024)
    return (__shfl_sync(0xffffffffu, mine, int(it >> 5)) >> (it & 31)) & 1u;
  const unsigned long long g = g0 + it;   // (ranges of more than 2^22 pixels per warp)
  return (__ldcg(gptr(d.lipsum) + (g >> 5)) >> (g & 31)) & 1u;
#else
  return true;
#endif
}

// Cluster mode: what the CTAs of a stream's cluster tell each other during the LIP pass (global
// memory; written before a cluster barrier, read after it).
struct LipBox {
  unsigned long long cnt[8], sig[8], endpos;
};

// R, rank, lb: the stream's cluster (R = 1: one CTA does it all). The stream side -- a scan over
// the bit string -- is the leader's; the mask side is split over the 32 R warps of the cluster.
static __device__ void dec_lip_pass(DecChunk& d, DecShared& S, int n_plane, int R = 1, int rank = 0,
                                    LipBox* lb = nullptr)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned long long K = S.klip;
  if (K == 0)
    return;   // the same on every CTA of the cluster
  if (rank == 0) {
  for (unsigned long long i = tid; i < (K + 31) / 32 + 1; i += kDecThreads) {
    gptr(d.sigarr)[i] = 0;
    gptr(d.signarr)[i] = 0;
  }
  block_sync();

  // stream side: tokenise until K pixels have been seen
  unsigned long long base_pos = S.pos, ord_base = 0;
  unsigned state_in = 0;
  while (ord_base < K) {
    const unsigned long long p = base_pos + 32ull * tid;
    const unsigned long long w64 = fetch64(d, p);
    const unsigned w = unsigned(w64);
    unsigned fA = lip_sim(w, 0), fB = lip_sim(w, 1);
    // inclusive scan of function composition inside the warp
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned pA = __shfl_up_sync(0xffffffffu, fA, o);
      const unsigned pB = __shfl_up_sync(0xffffffffu, fB, o);
      if (lane >= o) {
        unsigned nA, nB;
        lip_compose(pA, pB, fA, fB, nA, nB);
        fA = nA;
        fB = nB;
      }
    }
    if (lane == 31) {
      S.fA[warp] = fA;
      S.fB[warp] = fB;
    }
    block_sync();
    if (warp == 0) {
      unsigned gA = S.fA[lane], gB = S.fB[lane];
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned pA = __shfl_up_sync(0xffffffffu, gA, o);
        const unsigned pB = __shfl_up_sync(0xffffffffu, gB, o);
        if (lane >= o) {
          unsigned nA, nB;
          lip_compose(pA, pB, gA, gB, nA, nB);
          gA = nA;
          gB = nB;
        }
      }
      S.fA[lane] = gA;   // inclusive over warps
      S.fB[lane] = gB;
      if (lane == 31) {
        S.tile_fA = gA;
        S.tile_fB = gB;
      }
    }
    block_sync();
    // exclusive prefix of this thread = (warps before) then (lanes before)
    unsigned eA = __shfl_up_sync(0xffffffffu, fA, 1), eB = __shfl_up_sync(0xffffffffu, fB, 1);
    if (lane == 0) {
      eA = 0u;        // identity: state kept, no tokens
      eB = 1u;
    }
    if (warp > 0) {
      unsigned nA, nB;
      lip_compose(S.fA[warp - 1], S.fB[warp - 1], eA, eB, nA, nB);
      eA = nA;
      eB = nB;
    }
    const unsigned mine = state_in ? eB : eA;
    int s = int(mine & 1u);
    unsigned long long ord = ord_base + (mine >> 1);
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
      const unsigned b = (w >> i) & 1u;
      if (s == 0) {
        if (ord < K) {
          if (b) {
            atomicOr(&gptr(d.sigarr)[ord >> 5], 1u << (ord & 31));
            if ((w64 >> (i + 1)) & 1ull)
              atomicOr(&gptr(d.signarr)[ord >> 5], 1u << (ord & 31));
          }
          if (ord == K - 1)
            S.endpos = p + i + 1 + b;
        }
        ord++;
        s = int(b);
      }
      else
        s = 0;
    }
    const unsigned tile = state_in ? S.tile_fB : S.tile_fA;
    ord_base += tile >> 1;
    state_in = tile & 1u;
    base_pos += 32ull * kDecThreads;
    block_sync();
  }
  }   // rank == 0
  if (R > 1) {
    if (rank == 0 && tid == 0)
      __stcg(gptr(&lb->endpos), S.endpos);
    cluster_sync();   // the token arrays are complete (and visible: the barrier fences)
  }

  // mask side: the k-th LIP pixel in raster order owns token k. A lane takes four mask words at a
  // time (one 128-bit load; the mask arrays are 16-byte aligned and zero padded).
  const unsigned long long words = (d.n + 31) / 32;
  const unsigned long long nwarps = (unsigned long long)kDecWarps * R;
  const unsigned long long per_warp = ((words + nwarps - 1) / nwarps + 127) / 128 * 128;
  const unsigned long long w0 = per_warp * ((unsigned long long)rank * kDecWarps + warp),
                           w1 = min((words + 3) & ~3ull, w0 + per_warp);
  unsigned long long cnt = 0;
  const uint32_t summary = lip_summary_of(d, w0 >> 7, lane);   // (w0 is a multiple of 128)
  for (unsigned long long j0 = w0; j0 < w1; j0 += 128) {
    if (!lip_group_live(d, summary, w0 >> 7, (j0 - w0) >> 7))   // warp-uniform
      continue;
    const unsigned long long j = j0 + 4ull * lane;
    if (j >= w1)
      continue;
    // (.cg in cluster mode: other SMs set these bits, with atomics that live in L2)
    const uint4 m4 = lip_load(d, j);
    cnt += __popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w);
  }
  for (int o = 16; o; o >>= 1)
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  if (lane == 0)
    S.wtot[warp] = cnt;
  block_sync();
  if (warp == 0) {
    const unsigned long long v = S.wtot[lane];
    const unsigned long long inc = warp_incl_scan(v, lane);
    S.wtot[lane] = inc - v;
    if (lane == 31)
      S.wtot2[0] = inc;   // LIP pixels in this CTA's part of the mask
  }
  block_sync();
  unsigned long long run = S.wtot[warp];
  if (R > 1) {   // ... after those of the lower ranks
    if (tid == 0)
      __stcg(gptr(&lb->cnt[rank]), S.wtot2[0]);
    cluster_sync();
    for (int r2 = 0; r2 < rank; r2++)
      run += __ldcg(gptr(&lb->cnt[r2]));
  }
  unsigned long long nsig = 0;
  for (unsigned long long j0 = w0; j0 < w1; j0 += 128) {
    if (!lip_group_live(d, summary, w0 >> 7, (j0 - w0) >> 7))   // an empty group: no tokens, `run` unchanged
      continue;
    const unsigned long long j = j0 + 4ull * lane;
    uint4 m4 = make_uint4(0u, 0u, 0u, 0u);
    if (j < w1)
      m4 = lip_load(d, j);
    unsigned mw[4] = {m4.x, m4.y, m4.z, m4.w};
    const unsigned c4 = unsigned(__popc(m4.x) + __popc(m4.y) + __popc(m4.z) + __popc(m4.w));
    unsigned inc = c4;
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o)
        inc += t;
    }
    unsigned long long k = run + inc - c4;
    run += __shfl_sync(0xffffffffu, inc, 31);
    if (c4 == 0)
      continue;
    bool changed = false;
#pragma unroll
    for (int q = 0; q < 4; q++) {
      unsigned m = mw[q];
      const unsigned c = unsigned(__popc(m));
      if (c == 0)
        continue;
      // the c tokens of this word: bits [k, k + c) of the two result arrays
      const unsigned long long wi = k >> 5;
      const unsigned sh = unsigned(k & 31);
      k += c;
      unsigned sg = __funnelshift_r(lip_res_load(d, gptr(d.sigarr) + wi), lip_res_load(d, gptr(d.sigarr) + wi + 1), sh);
      if (c < 32)
        sg &= (1u << c) - 1u;
      if (sg == 0)
        continue;
      const unsigned sn = __funnelshift_r(lip_res_load(d, gptr(d.signarr) + wi), lip_res_load(d, gptr(d.signarr) + wi + 1), sh);
      unsigned keep = m;
      int t = 0;
      while (m) {
        const int bit = __ffs(m) - 1;
        m &= m - 1;
        if ((sg >> t) & 1u) {
          keep &= ~(1u << bit);
          gptr(d.pl)[(j + q) * 32 + bit] = uint8_t(n_plane | (((sn >> t) & 1u) ? 0 : 0x80));
        }
        t++;
      }
      mw[q] = keep;
      changed = true;
      nsig += __popc(sg);
    }
    if (changed)
      *reinterpret_cast<uint4*>(gptr(d.lip) + j) = make_uint4(mw[0], mw[1], mw[2], mw[3]);
  }
  for (int o = 16; o; o >>= 1)
    nsig += __shfl_xor_sync(0xffffffffu, nsig, o);
  block_sync();
  if (lane == 0)
    S.wtot[warp] = nsig;
  block_sync();
  if (R > 1) {
    if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
      unsigned long long t = 0;
      for (int i = 0; i < kDecWarps; i++)
        t += S.wtot[i];
      __stcg(gptr(&lb->sig[rank]), t);
    }
    cluster_sync();
    if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
      unsigned long long t = 0;
      for (int r2 = 0; r2 < R; r2++)
        t += __ldcg(gptr(&lb->sig[r2]));
      S.klip -= t;
      S.knew += t;
      S.pos = __ldcg(gptr(&lb->endpos));
    }
    block_sync();
    return;
  }
  if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
    unsigned long long t = 0;
    for (int i = 0; i < kDecWarps; i++)
      t += S.wtot[i];
    S.klip -= t;
    S.knew += t;
    S.pos = S.endpos;
  }
  block_sync();
}

// ---- end of a plane: where its refinement bits are (src/SPECK_INT.cpp:165-228, 359-469) ---------
// Called by thread 0 after the sorting pass of plane n. Returns false when decoding stops.
static __device__ bool dec_plane_end(DecChunk& d, DecShared& S, int n_plane)
{
  bool go = true;
  if (S.pos >= d.avail)
    go = false;   // the stream ended inside / right after the sorting pass: no refinement bits
  else {
    const unsigned long long nref = min(S.klsp, d.avail - S.pos);
    d.ref_base[n_plane] = S.pos;
    d.ref_cnt[n_plane] = nref;
    S.pos += nref;
    if (S.pos >= d.avail)
      go = false;
  }
  S.klsp += S.knew;
  S.knew = 0;
  return go;
}

// ---- LIS part: depth-first walk by one thread ---------------------------------------------------

constexpr int kDecMaxDepth = 48;

struct DecFrame {
  DChild kid[8];
  int nch, k, sigc;
  int all_tested;   // even the last child has its own significance bit (2D: last split of set I)
};

template <class T>
__device__ void dec_expand(DecChunk& d, const typename T::Data& tree, unsigned c, BitReader& br,
                           node_t root, int root_lis, DecFrame* stack, unsigned long long& klip,
                           unsigned long long& knew, int n_plane)
{
  int depth = 0;
  {
    const int nf = T::children(tree, d, c, root, root_lis, stack[0].kid);
    stack[0].nch = nf & 0xff;
    stack[0].all_tested = nf >> 8;
  }
  stack[0].k = 0;
  stack[0].sigc = 0;
  while (depth >= 0) {
    DecFrame& f = stack[depth];
    if (f.k == f.nch) {
      depth--;
      continue;
    }
    const DChild ch = f.kid[f.k];

    const bool need = f.sigc != 0 || f.k != f.nch - 1 || f.all_tested;
    f.k++;
    const unsigned sig = need ? br.get() : 1u;
    if (ch.pixel) {
      const unsigned long long i = ch.idx;
      if (sig) {
        const unsigned sgn = br.get();
        gptr(d.pl)[i] = uint8_t(n_plane | (sgn ? 0 : 0x80));
        knew++;
        f.sigc++;
      }
      else {
        gptr(d.lip)[i >> 5] |= 1u << (i & 31);
        lip_mark_group(d, i >> 5);
        klip++;
      }
    }
    else if (sig) {
      f.sigc++;
      if (depth + 1 >= kDecMaxDepth) {
        d.err |= 2u;
        return;
      }
      DecFrame& g = stack[depth + 1];
      const int nf = T::children(tree, d, c, ch.id, ch.lis, g.kid);
      g.nch = nf & 0xff;
      g.all_tested = nf >> 8;
      g.k = 0;
      g.sigc = 0;
      depth++;
    }
    else if (T::kHasI && ch.lis < 0)
      d.iset = ch.id;   // the set I stays outside the lists: it is tested after all of them
    else {
      const unsigned slot = gptr(d.lis_cnt)[ch.lis];
      if (gptr(d.lis_off)[ch.lis] + slot >= gptr(d.lis_off)[ch.lis + 1]) {
        d.err |= 1u;
        return;
      }
      gptr(d.lis)[gptr(d.lis_off)[ch.lis] + slot] = ch.id;
      gptr(d.lis_cnt)[ch.lis] = slot + 1;
    }
  }
}

template <class T>
__device__ void dec_lis_walk(DecChunk& d, const typename T::Data& tree, unsigned c, DecShared& S,
                             DecFrame* stack, int n_plane)
{
  BitReader br;
  br.init(d, S.pos);
  unsigned long long klip = S.klip, knew = S.knew;
  for (int lev = d.nlis - 1; lev >= 0; lev--) {
    const unsigned cnt = gptr(d.lis_cnt)[lev];
    if (cnt == 0)
      continue;
    node_t* list = gptr(d.lis) + gptr(d.lis_off)[lev];
    unsigned w = 0;
    for (unsigned i = 0; i < cnt; i++) {
      const node_t nd = list[i];
      if (br.get() == 0) {
        list[w++] = nd;
        continue;
      }
      dec_expand<T>(d, tree, c, br, nd, lev, stack, klip, knew, n_plane);
      if (d.err)
        return;
    }
    gptr(d.lis_cnt)[lev] = w;
  }
  if (T::kHasI && d.iset) {   // SPECK2D_INT::m_sorting_pass, third step (src/SPECK2D_INT.cpp:54-57)
    if (br.get()) {
      const node_t nd = d.iset;
      d.iset = 0;
      dec_expand<T>(d, tree, c, br, nd, -1, stack, klip, knew, n_plane);
      if (d.err)
        return;
    }
  }
  S.pos = br.pos;
  S.klip = klip;
  S.knew = knew;
}

// ---- the kernel ---------------------------------------------------------------------------------

template <class T>
__global__ void __launch_bounds__(kDecThreads) k_speck_decode(DecChunk* chunks, typename T::Data tree)
{
  __shared__ DecShared S;
  __shared__ DecFrame stack[kDecMaxDepth];
  __shared__ int s_go;
  const unsigned c = blockIdx.x;
  DecChunk& d = chunks[c];
  if (d.skip || d.planes == 0 || d.pow2 || d.kind != T::kKind)   // power-of-two trees: k_speck_decode_fast
    return;
  const int tid = threadIdx.x;
  if (tid == 0) {
    S.pos = 0;
    S.klip = 0;
    S.klsp = 0;
    S.knew = 0;
    S.endpos = 0;
    for (int l = 0; l < d.nlis; l++)
      gptr(d.lis_cnt)[l] = 0;
    d.iset = T::kHasI ? T::iset(tree, d, c) : 0;
    const int nr = T::num_roots(tree, d, c);
    for (int r = 0; r < nr; r++) {
      node_t nd;
      int lis;
      T::root(tree, d, c, r, nd, lis);
      gptr(d.lis)[gptr(d.lis_off)[lis] + gptr(d.lis_cnt)[lis]] = nd;
      gptr(d.lis_cnt)[lis]++;
    }
  }
  block_sync();
  int n = d.planes - 1;
  for (int bp = 0; bp < d.planes; bp++, n--) {
    dec_lip_pass(d, S, n);
    block_sync();   // every thread has read the LIP population before the walker changes it
    if (tid == 0) {
      dec_lis_walk<T>(d, tree, c, S, stack, n);
      s_go = (!d.err && dec_plane_end(d, S, n)) ? 1 : 0;
    }
    block_sync();
    if (!s_go)
      return;
  }
}

}  // namespace sperr_b200
