// Fast integer SPECK decoder for power-of-two trees (3D chunks whose extents are powers of two and
// 1D outlier arrays of power-of-two length): the LIS part of a bit-plane decoded by the whole CTA.
//
// Reference behaviour reproduced (bit-exactly): the sorting pass of SPECK3D_INT / SPECK1D_INT with
// the decoder's significance reads,
//   /root/reference/src/SPECK3D_INT.cpp:99-212, src/SPECK3D_INT_DEC.cpp:8-49,
//   /root/reference/src/SPECK1D_INT_DEC.cpp:12-125.
//
// How (not how the reference does it). A SPECK stream is a depth-first serialisation: where a set's
// bits start depends on everything before it, so a naive decoder is one long dependent chain. Two
// facts break the chain:
//   (1) The number of bits a *small* set consumes once it is known to be significant ("body") is a
//       function of the bits that follow it only. So body lengths can be computed speculatively for
//       EVERY bit position of a window at once: bodyA (sets whose children are single coefficients)
//       from two table look-ups, bodyB / bodyC (one / two levels up) by chaining 8 look-ups in the
//       table below. That is embarrassingly parallel work over the window.
//   (2) The LIS is visited list by list, deepest sets first, so a plane's LIS part starts with long
//       runs of small-set roots, each coded as [significance bit][body if 1]. With the speculative
//       lengths every position knows where the next token would start if a token started here;
//       pointer doubling from the known start of the window marks the real token starts, a prefix
//       count gives each token its root, and one thread per token then expands it.
// Larger sets are walked by one thread that only touches the top of the tree: whenever it reaches a
// bodyC-sized child it records a token and skips its body with one look-up.
// Tokens are expanded in parallel; the sets they leave behind are appended to the lists in token
// order (count -> scan -> copy), which is SPECK's list order.
#pragma once

#include "speck_dec.cuh"

namespace sperr_b200 {

constexpr int kFW = 8192;             // window: bit positions whose tokens are resolved per round
constexpr int kFLA = 1280;            // look-ahead bits staged behind the window
constexpr int kFBits = kFW + kFLA;
constexpr int kFBodyA = kFW + 1248;   // bodyA entries (bodyB needs up to +136 beyond its own range)
constexpr int kFBodyB = kFW + 1104;   // bodyB entries (bodyC needs up to +1096)
constexpr int kFNext = kFW + 1112;    // chain positions incl. the absorbing exits beyond the window
constexpr int kFTok = 1024;           // size-C tokens queued per round
constexpr int kFQ = 2048;             // size-B / size-A tokens in flight (a slice of 256 parents)
constexpr int kFSlice = kFQ / 8;
constexpr int kFMaxDepth = 32;
constexpr int kFRoots = 2048;         // roots of the walker's current list staged per round
// thread-block clusters: R CTAs decode R consecutive windows of one stream at the same time
constexpr int kFMaxR = 8;
constexpr int kFEntry = 1120;         // a chain enters a window at an offset below this (>= kFNext - kFW)
constexpr int kFScr = kFW + kFLA;     // sets one window can append to one list (each costs a bit)

// Mailbox of one stream's cluster in global memory. Every field is written by one CTA before a
// cluster barrier and read by the others after it.
struct ClusterBox {
  // decoder state, published by the leader (rank 0) after the parts it runs alone
  unsigned long long pos, klip, klsp, knew;
  unsigned cnt[kMaxLis];
  int iset, go, plane;
  unsigned err;
  // per rank, per round of a chain phase
  unsigned marks[kFMaxR], surv[kFMaxR], app[kFMaxR][2], errs[kFMaxR];
  unsigned long long dklip[kFMaxR], dknew[kFMaxR], newpos[kFMaxR];
  uint16_t exits[kFMaxR][kFEntry];   // window position a chain entering at offset e leaves the window at
  LipBox lip;
};

struct FastSmem {
  uint32_t bits[kFBits / 32 + 4];     // window of the stream, word aligned; bit q is at q + boff
  uint16_t bodyC[kFW + 8];
  uint8_t bodyB[kFBodyB + 8];
  uint8_t bodyA[kFBodyA + 8];
  // bits consumed by a set of the size class that is coded at this position with its significance
  // bit ([bit][body if 1]): length in the low bits, the significance bit on top
  uint8_t stepA[kFBodyA + 8];     // len | sig << 7
  uint16_t stepB[kFBodyB + 8];    // len | sig << 15
  // chain tables: [0], [1] ping-pong of the pointer doubling, [2] = 16 tokens ahead, [3] = 256 ahead
  uint16_t nxt[4][kFNext + 8];
  uint16_t anc256[40], anc16[33 * 16];
  unsigned n256, n16;
  int flag[3];
  uint32_t mark[kFW / 32];
  uint16_t T1[256], T2[2][256];       // pixel-set tables: sig(4) | sign(4) << 4 | bits << 8
  // token queues of one round: sets known to be significant, by size class (C: depth J-3, B: J-2,
  // A: J-1), as (node, window position of the first bit of the body)
  unsigned long long qc_node[kFTok];
  uint16_t qc_pos[kFTok];
  unsigned long long qb_node[kFQ];
  uint16_t qb_pos[kFQ];
  unsigned long long qa_node[kFQ];
  uint16_t qa_pos[kFQ];
  unsigned nqc, nqb, nqa;
  unsigned long long off[kMaxLis + 1];
  unsigned cnt[kMaxLis];
  unsigned long long wsum[kDecWarps];
  unsigned long long scan_total;
  unsigned boff;
  unsigned cutpos;
  // geometry
  int Dx, Dy, Dz, J;
  unsigned nx, ny;
  // 2D slices (SPECK2D_INT): children are coded in reverse raster order, a set's list index is its
  // depth, and the set I (everything outside the approximation band of transform level iset) is
  // tested after all lists; iset == 0: I is used up
  int is2d, iset;
  // walker (thread 0) state, kept here so that it can be suspended between windows
  int wk_j;                 // depth of the list being visited
  unsigned wk_i, wk_w, wk_cnt;
  int wk_depth;             // -1: between roots
  unsigned long long wk_node[kFMaxDepth];
  unsigned char wk_k[kFMaxDepth], wk_sig[kFMaxDepth];
  int wk_done;
  unsigned wk_q;            // window position of the walker
  unsigned long long win_base;   // absolute bit position of the window
  unsigned long long rs_node[kFRoots];   // staged roots: entries [rs_first, rs_first + rs_cnt) of the list
  uint32_t rs_gone[kFRoots / 32];        // staged roots that turned significant in this round
  unsigned rs_first, rs_cnt;
  int go;
  int plane;
  unsigned err;
  unsigned napp;             // 1D walker: sets logged for appending in this round (the log lives in nxt[])
  // cluster
  int R, rank;
  unsigned entry;            // offset at which the chain enters this CTA's window
  unsigned app_cnt[2];       // sets appended to the scratch of the two child lists in this round
  unsigned Tr, i0r, sumT, rstar, wsurv_r, sum_surv;
  unsigned long long prof[8];
};

// mailbox accesses bypass L1 (another SM wrote the data)
template <class T>
__device__ __forceinline__ T mb_ld(const T* p)
{
  return __ldcg(gptr(p));
}
template <class T>
__device__ __forceinline__ void mb_st(T* p, T v)
{
  __stcg(gptr(p), v);
}

#ifdef SPERR_EMUL
#define F_CLOCK() 0ll
#else
#define F_CLOCK() clock64()
#endif
// phase timer of thread 0 (see DecChunk::prof)
#define F_TIC() const long long f_t0 = F_CLOCK()
#define F_TOC(F, k)                                            \
  do {                                                         \
    if (threadIdx.x == 0)                                      \
      (F).prof[k] += (unsigned long long)(F_CLOCK() - f_t0);   \
  } while (0)

// List entries may have been written by another CTA of the cluster: read them from L2 (every cluster
// barrier is preceded by a __threadfence, so they are there). A lone CTA must NOT do that: its own
// plain stores are only guaranteed to be visible to the CTA's ordinary loads (measured on B200: an
// ld.global.cg right after a __syncthreads can miss a store of the same CTA).
__device__ __forceinline__ node_t f_list_load(const FastSmem& F, const node_t* p)
{
  return F.R > 1 ? __ldcg(p) : *p;
}

// ---- geometry of power-of-two trees -----------------------------------------------------------
// A set at depth j is the box (ix, iy, iz) of the grid with 2^min(j, D_a) cells along axis a; it is
// stored as (j << 32 | linear index), its list index is sum_a min(j, D_a).

__device__ __forceinline__ int f_lis(const FastSmem& F, int j)
{
  return F.is2d ? j : min(j, F.Dx) + min(j, F.Dy) + min(j, F.Dz);
}
__device__ __forceinline__ unsigned long long f_pack(const FastSmem& F, int j, unsigned ix, unsigned iy,
                                                     unsigned iz)
{
  const int bx = min(j, F.Dx), by = min(j, F.Dy);
  return ((unsigned long long)j << 32) | (unsigned long long)(ix | (iy << bx) | (iz << (bx + by)));
}
__device__ __forceinline__ void f_unpack(const FastSmem& F, unsigned long long nd, int& j, unsigned& ix,
                                         unsigned& iy, unsigned& iz)
{
  j = int(nd >> 32);
  const unsigned lin = unsigned(nd);
  const int bx = min(j, F.Dx), by = min(j, F.Dy);
  ix = lin & ((1u << bx) - 1u);
  iy = (lin >> bx) & ((1u << by) - 1u);
  iz = lin >> (bx + by);
}
__device__ __forceinline__ int f_nch(const FastSmem& F, int j)
{
  return 1 << (int(j < F.Dx) + int(j < F.Dy) + int(j < F.Dz));
}
// k-th child (x fastest) of a set at depth j
__device__ __forceinline__ void f_child(const FastSmem& F, int j, unsigned ix, unsigned iy, unsigned iz,
                                        int k, unsigned& jx, unsigned& jy, unsigned& jz)
{
  const int sx = j < F.Dx, sy = j < F.Dy, sz = j < F.Dz;
  if (F.is2d)   // BR, BL, TR, TL (/root/reference/src/SPECK2D_INT.cpp:109-148)
    k = (1 << (sx + sy + sz)) - 1 - k;
  jx = sx ? ix * 2 + (unsigned(k) & 1u) : ix;
  jy = sy ? iy * 2 + ((unsigned(k) >> sx) & 1u) : iy;
  jz = sz ? iz * 2 + ((unsigned(k) >> (sx + sy)) & 1u) : iz;
}

// ---- window access ------------------------------------------------------------------------------

__device__ __forceinline__ unsigned f_bit(const FastSmem& F, unsigned q)
{
  q += F.boff;
  return (F.bits[q >> 5] >> (q & 31)) & 1u;
}
// the 32 bits starting at window position q
__device__ __forceinline__ unsigned f_peek(const FastSmem& F, unsigned q)
{
  q += F.boff;
  const unsigned w = q >> 5, s = q & 31;
  return __funnelshift_r(F.bits[w], F.bits[w + 1], s);
}

// packed three-field block scan (21 bits per field); returns the exclusive prefix, total in F
__device__ __forceinline__ unsigned long long f_block_scan(FastSmem& F, unsigned long long v)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned long long inc = warp_incl_scan(v, lane);
  block_sync();   // previous users of wsum / scan_total are done
  if (lane == 31)
    F.wsum[warp] = inc;
  block_sync();
  if (warp == 0) {
    const unsigned long long w = F.wsum[lane];
    const unsigned long long winc = warp_incl_scan(w, lane);
    F.wsum[lane] = winc - w;
    if (lane == 31)
      F.scan_total = winc;
  }
  block_sync();
  return F.wsum[warp] + inc - v;
}

// ---- speculative body lengths -------------------------------------------------------------------

static __device__ void f_build_luts(FastSmem& F)
{
  for (int i = threadIdx.x; i < 256 * 3; i += kDecThreads) {
    const int which = i >> 8;   // 0: first four pixels; 1, 2: last four with no / some earlier hit
    const unsigned w = unsigned(i) & 255u;
    unsigned pos = 0, sig = 0, sgn = 0;
    int c = which == 2 ? 1 : 0;
    for (int k = 0; k < 4; k++) {
      const bool need = which == 0 || c != 0 || k != 3;
      const unsigned s = need ? (w >> pos++) & 1u : 1u;
      if (s) {
        sig |= 1u << k;
        sgn |= ((w >> pos++) & 1u) << k;
        c++;
      }
    }
    const uint16_t e = uint16_t(sig | (sgn << 4) | (pos << 8));
    if (which == 0)
      F.T1[w] = e;
    else
      F.T2[which - 1][w] = e;
  }
}

// bits consumed by the pixels of a significant bottom-level set starting at window position q;
// also returns the significance and sign masks (child order, x fastest)
__device__ __forceinline__ unsigned f_pixels(const FastSmem& F, unsigned q, int nch, unsigned& sigm,
                                             unsigned& sgnm)
{
  const unsigned u = f_peek(F, q);
  if (nch == 8) {
    const unsigned t1 = F.T1[u & 255u];
    const unsigned n1 = t1 >> 8;
    const unsigned t2 = F.T2[(t1 & 15u) ? 1 : 0][(u >> n1) & 255u];
    sigm = (t1 & 15u) | ((t2 & 15u) << 4);
    sgnm = ((t1 >> 4) & 15u) | (((t2 >> 4) & 15u) << 4);
    return n1 + (t2 >> 8);
  }
  unsigned pos = 0;
  int c = 0;
  sigm = sgnm = 0;
  for (int k = 0; k < nch; k++) {
    const bool need = c != 0 || k != nch - 1;
    const unsigned s = need ? (u >> pos++) & 1u : 1u;
    if (s) {
      sigm |= 1u << k;
      sgnm |= ((u >> pos++) & 1u) << k;
      c++;
    }
  }
  return pos;
}

// Stages the stream window that starts at absolute bit position `base` and fills the body tables
// up to `kinds` (0: bodyA only, 1: + bodyB, 2: + bodyC).
static __device__ void f_build_window_impl(const DecChunk& d, FastSmem& F, unsigned long long base,
                                           int kinds, bool after_one);
// after_one: the top table is only needed at positions that follow a set bit (roots of a list are
// coded as [significance bit][body if 1])
static __device__ void f_build_window(const DecChunk& d, FastSmem& F, unsigned long long base, int kinds,
                                      bool after_one)
{
  F_TIC();
  f_build_window_impl(d, F, base, kinds, after_one);
  block_sync();
  F_TOC(F, 1);
  if (threadIdx.x == 0)
    F.prof[6]++;
}
static __device__ void f_build_window_impl(const DecChunk& d, FastSmem& F, unsigned long long base,
                                           int kinds, bool after_one)
{
  const int tid = threadIdx.x;
  block_sync();   // every reader of the previous window is done
  const unsigned long long w0 = base >> 5;
  const int nbits = kinds == 0 ? kFW + 64 : (kinds == 1 ? kFW + 256 : kFBits);
  for (int i = tid; i < nbits / 32 + 3; i += kDecThreads) {
    const unsigned long long w = w0 + i;
    F.bits[i] = w < d.stage_words ? gptr(d.bits)[w] : 0u;
  }
  if (tid == 0)
    F.boff = unsigned(base & 31);
  block_sync();
  const int J = F.J;
  const int nchA = f_nch(F, J - 1);
  const int nA = kinds == 0 ? kFW + 8 : (kinds == 1 ? kFW + 144 : kFBodyA);
  // constant trip counts, fully unrolled: the table look-ups of a thread's positions are independent
  // chains of shared-memory loads and overlap instead of queueing behind one another
#pragma unroll
  for (int it = 0; it < (kFBodyA + kDecThreads - 1) / kDecThreads; it++) {
    const int q = tid + it * kDecThreads;
    if (q < nA) {
      unsigned sm, gm;
      F.bodyA[q] = uint8_t(f_pixels(F, q, nchA, sm, gm));
    }
  }
  block_sync();
  for (int q = tid; q < nA - 1; q += kDecThreads)
    F.stepA[q] = f_bit(F, q) ? uint8_t((1u + F.bodyA[q + 1]) | 0x80u) : uint8_t(1);
  if (kinds == 0)
    return;
  block_sync();
  const int nchB = f_nch(F, J - 2);
  const int nB = kinds == 1 ? kFW + 8 : kFBodyB;
#pragma unroll
  for (int it = 0; it < (kFBodyB + kDecThreads - 1) / kDecThreads; it++) {
    const int q = tid + it * kDecThreads;
    if (q >= nB || (kinds == 1 && after_one && q > 0 && !f_bit(F, q - 1)))
      continue;
    unsigned pos = q, c = 0;
    for (int k = 0; k < nchB - 1; k++) {   // all but the last child carry a significance bit
      const unsigned s = F.stepA[pos];
      c |= s;
      pos += s & 127u;
    }
    pos += (c & 0x80u) ? (F.stepA[pos] & 127u) : F.bodyA[pos];   // the last one may be inferred
    F.bodyB[q] = uint8_t(pos - q);
  }
  if (kinds == 1 && after_one) {
    block_sync();
    for (int q = tid; q < nB - 1; q += kDecThreads)
      F.stepB[q] = f_bit(F, q) ? uint16_t((1u + F.bodyB[q + 1]) | 0x8000u) : uint16_t(1);
    return;
  }
  if (kinds == 1)
    return;
  block_sync();
  for (int q = tid; q < nB - 1; q += kDecThreads)
    F.stepB[q] = f_bit(F, q) ? uint16_t((1u + F.bodyB[q + 1]) | 0x8000u) : uint16_t(1);
  block_sync();
  const int nchC = f_nch(F, J - 3);
#pragma unroll
  for (int it = 0; it < (kFW + 8 + kDecThreads - 1) / kDecThreads; it++) {
    const int q = tid + it * kDecThreads;
    if (q >= kFW + 8 || (after_one && q > 0 && !f_bit(F, q - 1)))
      continue;
    unsigned pos = q, c = 0;
    for (int k = 0; k < nchC - 1; k++) {
      const unsigned s = F.stepB[pos];
      c |= s;
      pos += s & 0x7fffu;
    }
    pos += (c & 0x8000u) ? (F.stepB[pos] & 0x7fffu) : F.bodyB[pos];
    F.bodyC[q] = uint16_t(pos - q);
  }
}

// ---- token expansion ------------------------------------------------------------------------------
// A round's tokens are expanded level by level: one thread per size-C token finds its children with
// the bodyB table (significant ones become size-B tokens, the others join list B), the same one
// level down with bodyA, and finally one thread per size-A token decodes its pixels. Queue slots and
// list positions come from block scans, so both keep depth-first (= stream) order.

// pixels of a significant bottom-level set (depth J - 1) whose body starts at window position q
static __device__ void f_expand_pixels(const DecChunk& d, const FastSmem& F, unsigned long long nd,
                                       unsigned q, int n_plane, unsigned& nlip, unsigned& nsig)
{
  int j;
  unsigned ix, iy, iz;
  f_unpack(F, nd, j, ix, iy, iz);
  const int sx = j < F.Dx, sy = j < F.Dy, sz = j < F.Dz;
  const int nch = 1 << (sx + sy + sz);
  unsigned sigm, sgnm;
  f_pixels(F, q, nch, sigm, sgnm);
  if (F.is2d) {   // masks are in coding order = reverse raster order
    sigm = __brev(sigm) >> (32 - nch);
    sgnm = __brev(sgnm) >> (32 - nch);
  }
  nsig += unsigned(__popc(sigm));
  nlip += unsigned(nch - __popc(sigm));
  const unsigned x0 = sx ? ix * 2 : ix, y0 = sy ? iy * 2 : iy, z0 = sz ? iz * 2 : iz;
  const unsigned long long nx = F.nx, nxy = (unsigned long long)F.nx * F.ny;
  const unsigned long long base = (unsigned long long)z0 * nxy + (unsigned long long)y0 * nx + x0;
  // pixels that differ in x only are neighbours in the LIP mask: one update per row
  const int per_row = sx ? 2 : 1;
  const int rows = nch / per_row;
  const unsigned rm = sx ? 3u : 1u;
  for (int r = 0; r < rows; r++) {
    const unsigned cy = sy ? (unsigned(r) & 1u) : 0u;
    const unsigned cz = sz ? ((unsigned(r) >> sy) & 1u) : 0u;
    const unsigned long long i = base + cy * nx + cz * nxy;
    const unsigned s = (sigm >> (r * per_row)) & rm, g = (sgnm >> (r * per_row)) & rm;
    if (s & 1u)
      gptr(d.pl)[i] = uint8_t(n_plane | ((g & 1u) ? 0 : 0x80));
    if (s & 2u)
      gptr(d.pl)[i + 1] = uint8_t(n_plane | ((g & 2u) ? 0 : 0x80));
    if (rm & ~s) {   // x0 is even when a row has two pixels: both bits are in the same word
      atomicOr(&gptr(d.lip)[i >> 5], (rm & ~s) << (i & 31));
      lip_mark_group(d, i >> 5);
    }
  }
}

// One level of expansion: `np` parent tokens (pn, pp) at depth j; significant children go to the
// child queue (cn, cp, *ncq), the others to the list of depth j + 1. CHILD_TAB: body table of the
// children (bodyB for size-C parents, bodyA for size-B parents).
// CL: appends go to this CTA's scratch of the child list (slot 0: depth J - 2, 1: depth J - 1); the
// cluster copies them to the lists in rank order at the end of the round.
template <class Tab, bool CL>
static __device__ void f_expand_level(DecChunk& d, FastSmem& F, int j, const unsigned long long* pn,
                                      const uint16_t* pp, unsigned np, const Tab* child_body,
                                      unsigned long long* cn, uint16_t* cp, unsigned* ncq)
{
  const int tid = threadIdx.x;
  const int nch = f_nch(F, j);
  const int cl = f_lis(F, j + 1);
  // parents beyond the thread count are handled in further sweeps (np <= kFTok or kFSlice here)
  for (unsigned p0 = 0; p0 < np; p0 += kDecThreads) {
    const unsigned p = p0 + tid;
    unsigned long long kid[8];
    uint16_t kpos[8];
    unsigned sigmask = 0;
    int nk = 0;
    if (p < np) {
      int jj;
      unsigned ix, iy, iz;
      f_unpack(F, pn[p], jj, ix, iy, iz);
      unsigned q = pp[p];
      int c = 0;
      for (int k = 0; k < nch; k++) {
        const bool need = c != 0 || k != nch - 1;
        const unsigned s = need ? f_bit(F, q++) : 1u;
        unsigned jx, jy, jz;
        f_child(F, j, ix, iy, iz, k, jx, jy, jz);
        kid[k] = f_pack(F, j + 1, jx, jy, jz);
        kpos[k] = uint16_t(q);
        if (s) {
          c = 1;
          sigmask |= 1u << k;
          q += child_body[q];
        }
      }
      nk = nch;
    }
    const unsigned nsig = unsigned(__popc(sigmask)), nins = unsigned(nk) - nsig;
    const unsigned long long ex = f_block_scan(F, (unsigned long long)nsig | ((unsigned long long)nins << 21));
    const unsigned long long tot = F.scan_total;
    const unsigned exs = unsigned(ex) & 0x1fffffu, exi = unsigned(ex >> 21) & 0x1fffffu;
    const unsigned tots = unsigned(tot) & 0x1fffffu, toti = unsigned(tot >> 21) & 0x1fffffu;
    const unsigned qbase = *ncq;
    const int slot = j + 3 - F.J;   // CL only: children at depth J - 2 -> 0, J - 1 -> 1
    const unsigned long long lbase =
        CL ? ((unsigned long long)(F.rank * 2 + slot) * kFScr + F.app_cnt[slot]) : F.off[cl] + F.cnt[cl];
    const bool ok = (CL ? F.app_cnt[slot] + toti <= unsigned(kFScr) : lbase + toti <= F.off[cl + 1]) &&
                    qbase + tots <= unsigned(kFQ);
    if (ok) {
      unsigned a = qbase + exs;
      unsigned long long b = lbase + exi;
      node_t* const dst = CL ? gptr(d.scr) : gptr(d.lis);
      for (int k = 0; k < nk; k++) {
        if ((sigmask >> k) & 1u) {
          cn[a] = kid[k];
          cp[a] = kpos[k];
          a++;
        }
        else
          dst[b++] = kid[k];
      }
    }
    block_sync();
    if (tid == 0) {
      if (!ok)
        F.err |= 1u;
      *ncq = qbase + tots;
      if (CL)
        F.app_cnt[slot] += toti;
      else
        F.cnt[cl] += toti;
    }
    block_sync();
  }
}

// size-A tokens: decode the pixels
static __device__ void f_expand_A(DecChunk& d, DecShared& S, FastSmem& F, const unsigned long long* an,
                                  const uint16_t* ap, unsigned na, int n_plane)
{
  unsigned nlip = 0, nsig = 0;
  for (unsigned p = threadIdx.x; p < na; p += kDecThreads)
    f_expand_pixels(d, F, an[p], ap[p], n_plane, nlip, nsig);
  f_block_scan(F, (unsigned long long)nlip | ((unsigned long long)nsig << 32));
  if (threadIdx.x == 0) {
    S.klip += F.scan_total & 0xffffffffull;
    S.knew += F.scan_total >> 32;
  }
  block_sync();
}

// Expands the tokens of a round. kind: size class of the tokens in the entry queue
// (2: qc, 1: qb, 0: qa).
template <bool CL>
static __device__ void f_expand_round(DecChunk& d, DecShared& S, FastSmem& F, int kind, int n_plane)
{
  const int J = F.J;
  if (kind == 0) {
    f_expand_A(d, S, F, F.qa_node, F.qa_pos, F.nqa, n_plane);
    return;
  }
  if (kind == 1) {
    // the entry queue may hold up to kFQ size-B tokens: slices of kFSlice parents
    const unsigned nb = F.nqb;
    for (unsigned b0 = 0; b0 < nb; b0 += kFSlice) {
      if (threadIdx.x == 0)
        F.nqa = 0;
      block_sync();
      f_expand_level<uint8_t, CL>(d, F, J - 2, F.qb_node + b0, F.qb_pos + b0, min(unsigned(kFSlice), nb - b0),
                                  F.bodyA, F.qa_node, F.qa_pos, &F.nqa);
      if (F.err)
        return;
      f_expand_A(d, S, F, F.qa_node, F.qa_pos, F.nqa, n_plane);
    }
    return;
  }
  const unsigned nc = F.nqc;
  for (unsigned c0 = 0; c0 < nc; c0 += kFSlice) {
    if (threadIdx.x == 0)
      F.nqb = 0;
    block_sync();
    f_expand_level<uint8_t, CL>(d, F, J - 3, F.qc_node + c0, F.qc_pos + c0, min(unsigned(kFSlice), nc - c0),
                                F.bodyB, F.qb_node, F.qb_pos, &F.nqb);
    if (F.err)
      return;
    const unsigned nb = F.nqb;
    for (unsigned b0 = 0; b0 < nb; b0 += kFSlice) {
      if (threadIdx.x == 0)
        F.nqa = 0;
      block_sync();
      f_expand_level<uint8_t, CL>(d, F, J - 2, F.qb_node + b0, F.qb_pos + b0, min(unsigned(kFSlice), nb - b0),
                                  F.bodyA, F.qa_node, F.qa_pos, &F.nqa);
      if (F.err)
        return;
      f_expand_A(d, S, F, F.qa_node, F.qa_pos, F.nqa, n_plane);
    }
  }
}

// ---- phase A: the lists of the three smallest set sizes, by token chains -------------------------
//
// One round resolves R consecutive windows of the stream, one per CTA of the cluster (R = 1: one
// window). Everything that does not depend on where the token chain enters a window is computed by
// all CTAs at once: the body tables, the token length of every position, and -- by pointer
// doubling -- where a chain entering at any offset leaves the window. Then the entry offsets are
// handed down the ranks (one table look-up each), every CTA marks the token starts of its window
// from its true entry (hierarchically: 256-token and 16-token jump tables kept from the doubling),
// and the windows' tokens are mapped to the roots of the list, expanded, and their results appended
// in rank order, which is stream order.

__device__ __forceinline__ unsigned f_token_len(const FastSmem& F, int kind, unsigned q)
{
  if (kind == 0)
    return F.stepA[q] & 127u;
  if (kind == 1)
    return F.stepB[q] & 0x7fffu;
  return f_bit(F, q) ? 1u + F.bodyC[q + 1] : 1u;
}

template <bool CL>
static __device__ void f_list_by_chain(DecChunk& d, DecShared& S, FastSmem& F, int kind, int n_plane)
{
  const int tid = threadIdx.x;
  const int R = CL ? F.R : 1, rank = CL ? F.rank : 0;
  const int j = F.J - 1 - kind;
  const int lis = f_lis(F, j);
  const unsigned m = F.cnt[lis];
  if (m == 0)
    return;
  ClusterBox* const box = d.box;
  node_t* const list = gptr(d.lis) + F.off[lis];
  unsigned long long* const qn = kind == 0 ? F.qa_node : (kind == 1 ? F.qb_node : F.qc_node);
  uint16_t* const qp = kind == 0 ? F.qa_pos : (kind == 1 ? F.qb_pos : F.qc_pos);
  const unsigned qcap = kind == 2 ? unsigned(kFTok) : unsigned(kFQ);
  const int n_entry = CL ? kFEntry : 1;
  unsigned i0 = 0, wsurv = 0;
  while (i0 < m) {
    const unsigned long long round_base = S.pos;
    f_build_window(d, F, round_base + (unsigned long long)rank * kFW, kind, true);
    const long long f_tc = F_CLOCK();
    // token length of every window position; positions beyond the window absorb
    for (int q = tid; q < kFNext; q += kDecThreads)
      F.nxt[0][q] = uint16_t(q < kFW ? q + f_token_len(F, kind, q) : q);
    for (int i = tid; i < kFW / 32; i += kDecThreads)
      F.mark[i] = 0u;
    if (tid == 0) {
      F.flag[0] = 0;
      F.app_cnt[0] = F.app_cnt[1] = 0;
    }
    block_sync();
    // pointer doubling: table r + 1 = 2^(r+1) tokens ahead; until every entry offset has left
    int src = 0, have4 = 0, have8 = 0;
    for (int r = 0; r < 14; r++) {
      const int dst = r == 3 ? 2 : (r == 7 ? 3 : (src == 0 ? 1 : 0));
      const uint16_t* cn = F.nxt[src];
      uint16_t* nn = F.nxt[dst];
      if (tid == 0)
        F.flag[(r + 1) % 3] = 0;   // (round r + 1 resets the flag of round r + 2: nobody reads it any more)
#pragma unroll
      for (int it = 0; it < (kFNext + kDecThreads - 1) / kDecThreads; it++) {
        const int q = tid + it * kDecThreads;
        if (q < kFNext) {
          const uint16_t v = q < kFW ? cn[cn[q]] : uint16_t(q);
          nn[q] = v;
          if (q < n_entry && v < kFW)
            F.flag[r % 3] = 1;
        }
      }
      block_sync();
      src = dst;
      have4 |= r == 3;
      have8 |= r == 7;
      if (!F.flag[r % 3])
        break;
    }
    const uint16_t* const fin = F.nxt[src];
    // where does the chain enter my window?
    unsigned entry = 0;
    if (CL) {
      for (int e = tid; e < kFEntry; e += kDecThreads)
        mb_st(&box->exits[rank][e], fin[e]);
      cluster_sync();
      uint16_t* const xtab = reinterpret_cast<uint16_t*>(F.qa_node);   // idle until the expansion
      for (int e = tid; e < rank * kFEntry; e += kDecThreads)
        xtab[e] = mb_ld(&box->exits[0][0] + e);
      block_sync();
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        unsigned e = 0;
        for (int s2 = 0; s2 < rank; s2++)
          e = unsigned(xtab[s2 * kFEntry + e]) - unsigned(kFW);
        F.entry = e;
      }
      block_sync();
      entry = F.entry;
    }
    const unsigned exitpos = fin[entry];
    // token starts of my window: anchors every 256 tokens, then every 16, then single steps
    if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
      unsigned n = 0;
      if (have8) {
        unsigned p2 = entry;
        while (p2 < unsigned(kFW) && n < 33u) {
          F.anc256[n++] = uint16_t(p2);
          p2 = F.nxt[3][p2];
        }
      }
      else
        F.anc256[n++] = uint16_t(entry);
      F.n256 = n;
    }
    block_sync();
    if (have4) {
      if (tid < int(F.n256)) {
        unsigned p2 = F.anc256[tid];
        for (int u = 0; u < 16; u++) {
          F.anc16[tid * 16 + u] = uint16_t(p2 < unsigned(kFW) ? p2 : 0xFFFFu);
          if (p2 < unsigned(kFW))
            p2 = F.nxt[2][p2];
        }
      }
      if (tid == 0)
        F.n16 = F.n256 * 16;
    }
    else if (tid == 0) {
      F.anc16[0] = F.anc256[0];
      F.n16 = 1;
    }
    block_sync();
    if (tid < int(F.n16)) {
      unsigned p2 = F.anc16[tid];
      for (int u = 0; u < 16 && p2 < unsigned(kFW); u++) {
        atomicOr(&F.mark[p2 >> 5], 1u << (p2 & 31));
        p2 += f_token_len(F, kind, p2);
      }
    }
    block_sync();
    // rank of my marks = index of the root they belong to
    const unsigned mb = (F.mark[tid >> 2] >> ((tid & 3) * 8)) & 0xffu;
    unsigned long long ex = f_block_scan(F, (unsigned long long)__popc(mb));
    const unsigned total_marks = unsigned(F.scan_total);
    // which roots are mine: tokens of the windows before mine come first
    unsigned T, i0r, sumT, rstar = 0;
    if (CL) {
      if (tid == 0)
        mb_st(&box->marks[rank], total_marks);
      cluster_sync();
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        unsigned at = i0, sum = 0, mine_i0 = 0, mine_T = 0, last = 0;
        for (int s2 = 0; s2 < R; s2++) {
          const unsigned mk = mb_ld(&box->marks[s2]);
          const unsigned t2 = min(mk, m - at);
          if (s2 == rank) {
            mine_i0 = at;
            mine_T = t2;
          }
          if (t2 > 0)
            last = unsigned(s2);
          at += t2;
          sum += t2;
        }
        F.i0r = mine_i0;
        F.Tr = mine_T;
        F.sumT = sum;
        F.rstar = last;
      }
      block_sync();
      T = F.Tr;
      i0r = F.i0r;
      sumT = F.sumT;
      rstar = F.rstar;
    }
    else {
      T = min(total_marks, m - i0);
      i0r = i0;
      sumT = T;
    }
    // my tokens: survivors keep their place in the list, significant ones are queued
    node_t surv[8], sigs[8];
    uint16_t sigq[8];
    int ns = 0, ng = 0;
    {
      unsigned t = mb, r = unsigned(ex);
      while (t) {
        const int b = __ffs(t) - 1;
        t &= t - 1;
        const unsigned q = tid * 8 + b;
        if (r < T) {
          const node_t nd = f_list_load(F, &list[i0r + r]);
          if (f_bit(F, q)) {
            sigs[ng] = nd;
            sigq[ng++] = uint16_t(q + 1);
          }
          else
            surv[ns++] = nd;
        }
        else if (r == T)
          F.cutpos = q;   // first position after this list's last token of the window
        r++;
      }
    }
    // barriers inside: every root of the window has been read before the survivors are written
    ex = f_block_scan(F, (unsigned long long)ns | ((unsigned long long)ng << 32));
    const unsigned tot_surv = unsigned(F.scan_total & 0xffffffffull);
    const unsigned tot_sig = unsigned(F.scan_total >> 32);
    unsigned wsurv_r = wsurv, sum_surv = tot_surv;
    if (CL) {
      if (tid == 0)
        mb_st(&box->surv[rank], tot_surv);
      cluster_sync();   // ... in every window of the round
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        unsigned before = 0, sum = 0;
        for (int s2 = 0; s2 < R; s2++) {
          const unsigned v = mb_ld(&box->surv[s2]);
          if (s2 < rank)
            before += v;
          sum += v;
        }
        F.wsurv_r = wsurv + before;
        F.sum_surv = sum;
      }
      block_sync();
      wsurv_r = F.wsurv_r;
      sum_surv = F.sum_surv;
    }
    for (int s2 = 0; s2 < ns; s2++)
      list[wsurv_r + unsigned(ex & 0xffffffffull) + s2] = surv[s2];
    const long long f_te = F_CLOCK();
    if (tid == 0)
      F.prof[2] += (unsigned long long)(f_te - f_tc);
    const unsigned long long klip0 = S.klip, knew0 = S.knew;
    // the significant roots enter the queue of their size class, qcap at a time
    const unsigned myq = unsigned(ex >> 32);
    for (unsigned g0 = 0; g0 < tot_sig && !F.err; g0 += qcap) {
      for (int g = 0; g < ng; g++) {
        const unsigned slot = myq + g;
        if (slot >= g0 && slot < g0 + qcap) {
          qn[slot - g0] = sigs[g];
          qp[slot - g0] = sigq[g];
        }
      }
      if (tid == 0) {
        const unsigned cntq = min(qcap, tot_sig - g0);
        if (kind == 0) F.nqa = cntq;
        else if (kind == 1) F.nqb = cntq;
        else F.nqc = cntq;
      }
      block_sync();
      f_expand_round<CL>(d, S, F, kind, n_plane);
      block_sync();
    }
    if (tid == 0)
      F.prof[3] += (unsigned long long)(F_CLOCK() - f_te);
    if (!CL) {
      if (F.err)
        return;
      if (tid == 0)
        S.pos += T == total_marks ? exitpos : F.cutpos;
    }
    else {
      // end of the round: scratch appends to the lists in rank order, counters and position agreed
      if (tid == 0) {
        mb_st(&box->app[rank][0], F.app_cnt[0]);
        mb_st(&box->app[rank][1], F.app_cnt[1]);
        mb_st(&box->dklip[rank], S.klip - klip0);
        mb_st(&box->dknew[rank], S.knew - knew0);
        mb_st(&box->errs[rank], F.err);
        if (unsigned(rank) == rstar)
          mb_st(&box->newpos[rank], round_base + (unsigned long long)rank * kFW +
                                        (T == total_marks ? exitpos : F.cutpos));
      }
      cluster_sync();
      for (int a = 0; a < 2; a++) {
        unsigned before = 0, sum = 0;
        for (int s2 = 0; s2 < R; s2++) {
          const unsigned v = mb_ld(&box->app[s2][a]);
          if (s2 < rank)
            before += v;
          sum += v;
        }
        if (sum == 0)
          continue;   // the same on every thread of every rank
        const int cl = f_lis(F, F.J - 2 + a);
        const unsigned long long at = F.off[cl] + F.cnt[cl];
        const bool fits = at + sum <= F.off[cl + 1];
        const unsigned mine = F.app_cnt[a];
        if (fits) {
          const node_t* const from = gptr(d.scr) + (unsigned long long)(rank * 2 + a) * kFScr;
          for (unsigned t = tid; t < mine; t += kDecThreads)
            gptr(d.lis)[at + before + t] = from[t];
        }
        block_sync();   // every thread has read F.cnt[cl]
        if (tid == 0) {
          if (!fits)
            F.err |= 1u;
          F.cnt[cl] += sum;
        }
      }
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        unsigned long long dl = 0, dn = 0;
        unsigned e = 0;
        for (int s2 = 0; s2 < R; s2++) {
          dl += mb_ld(&box->dklip[s2]);
          dn += mb_ld(&box->dknew[s2]);
          e |= mb_ld(&box->errs[s2]);
        }
        S.klip = klip0 + dl;
        S.knew = knew0 + dn;
        S.pos = mb_ld(&box->newpos[rstar]);
        F.err |= e;
      }
      block_sync();
      if (F.err)
        return;   // the same on every rank
    }
    i0 += sumT;
    wsurv += sum_surv;
    block_sync();
  }
  if (tid == 0)
    F.cnt[lis] = wsurv;
  block_sync();
}

// ---- phase B: larger sets, one thread walks the top of the tree ----------------------------------

// Runs until the list is exhausted, the window or the token queue is used up, or the staged roots of
// the current list have all been visited. Roots that turn significant are only MARKED (rs_gone); the
// survivors are compacted into the list by the whole CTA afterwards (f_compact_roots), so a run of
// insignificant roots costs the walker one look at 32 bits, not one store per root.
// MODE 0: 3D chunk, 1: 1D outlier array (binary tree), 2: 2D slice (reverse child order, set I).
template <int MODE>
static __device__ void f_walk(DecChunk& d, DecShared& S, FastSmem& F)
{
  // geometry in registers: the helpers below are the f_* functions without the shared-memory reads
  const int Dx = F.Dx, Dy = MODE == 1 ? 0 : F.Dy, Dz = MODE == 0 ? F.Dz : 0;
  const unsigned boff = F.boff;
  constexpr int kIFrame = 0x100;   // frames of the set I (2D) carry depth kIFrame | part_level
  auto bit_at = [&](unsigned qq) {
    qq += boff;
    return (F.bits[qq >> 5] >> (qq & 31)) & 1u;
  };
  auto peek_at = [&](unsigned qq) {
    qq += boff;
    return __funnelshift_r(F.bits[qq >> 5], F.bits[(qq >> 5) + 1], qq & 31);
  };
  auto nch_of = [&](int jj) {
    if (MODE == 2 && (jj & kIFrame))
      return ((jj & 0xff) > 1) ? 4 : 3;
    if (MODE == 1)
      return 2;
    return 1 << (int(jj < Dx) + int(jj < Dy) + int(jj < Dz));
  };
  auto lis_of = [&](int jj) { return MODE == 0 ? min(jj, Dx) + min(jj, Dy) + min(jj, Dz) : jj; };
  auto pack = [&](int jj, unsigned x, unsigned y, unsigned z) {
    if (MODE == 1)
      return ((unsigned long long)jj << 32) | x;
    const int bx = min(jj, Dx), by = min(jj, Dy);
    return ((unsigned long long)jj << 32) | (unsigned long long)(x | (y << bx) | (z << (bx + by)));
  };
  auto unpack = [&](unsigned long long nd, int& jj, unsigned& x, unsigned& y, unsigned& z) {
    jj = int(nd >> 32);
    const unsigned lin = unsigned(nd);
    if (MODE == 1) {
      x = lin; y = 0; z = 0;
      return;
    }
    const int bx = min(jj, Dx), by = min(jj, Dy);
    x = lin & ((1u << bx) - 1u);
    y = (lin >> bx) & ((1u << by) - 1u);
    z = MODE == 0 ? lin >> (bx + by) : 0u;
  };

  unsigned q = F.wk_q;       // window position
  unsigned ntok = 0;
  int depth = F.wk_depth;
  unsigned i = F.wk_i;
  const unsigned cnt = F.wk_cnt;
  const unsigned i_end = F.rs_first + F.rs_cnt;
  const int jC = F.J - 3;
  // the frame being expanded lives in registers; outer frames are parked in shared memory
  int j = 0, k = 0, sg = 0, nch = 0;
  unsigned ix = 0, iy = 0, iz = 0;
  if (depth >= 0) {
    unpack(F.wk_node[depth], j, ix, iy, iz);
    k = F.wk_k[depth];
    sg = F.wk_sig[depth];
    nch = nch_of(j);
  }
  for (;;) {
    if (depth < 0) {
      if (i == cnt)
        break;   // list finished (the caller moves on to the next one)
      if (q + 1 >= unsigned(kFW) || i >= i_end)
        break;
      // roots coded as 0 stay in the list: skip a whole run of them at once
      const unsigned room = min(min(32u, i_end - i), unsigned(kFW) - 1u - q);
      unsigned u = peek_at(q);
      if (room < 32)
        u &= (1u << room) - 1u;
      const unsigned run = u ? unsigned(__ffs(u) - 1) : room;
      i += run;
      q += run;
      if (run == room)
        continue;
      const unsigned r = i - F.rs_first;
      F.rs_gone[r >> 5] |= 1u << (r & 31);
      const node_t nd = F.rs_node[r];
      i++;
      q++;
      depth = 0;
      unpack(nd, j, ix, iy, iz);
      k = 0;
      sg = 0;
      nch = nch_of(j);
      if (MODE == 2 && (j & kIFrame))
        F.iset = 0;   // I is being split; what is left of it is recorded below
      continue;
    }
    if (k == nch) {   // pop
      depth--;
      if (depth >= 0) {
        unpack(F.wk_node[depth], j, ix, iy, iz);
        k = F.wk_k[depth];
        sg = F.wk_sig[depth];
        nch = nch_of(j);
      }
      continue;
    }
    if (q + 1 >= unsigned(kFW) || ntok >= unsigned(kFTok))
      break;
    const bool is_i = MODE == 2 && (j & kIFrame) != 0;
    const int il = j & 0xff;
    // SPECK2D_INT::m_code_I (src/SPECK2D_INT.cpp:82-95): the three S sets are always tested; when
    // I's last level is split nothing follows them
    const bool need = sg != 0 || k != nch - 1 || (is_i && il == 1);
    const unsigned s = need ? bit_at(q++) : 1u;
    unsigned jx, jy, jz;
    int cj;
    bool child_i = false;
    if (!is_i) {
      const int sx = j < Dx, sy = j < Dy, sz = j < Dz;
      const int kk = MODE == 2 ? nch - 1 - k : k;   // 2D: BR, BL, TR, TL (src/SPECK2D_INT.cpp:109-148)
      jx = sx ? ix * 2 + (unsigned(kk) & 1u) : ix;
      jy = sy ? iy * 2 + ((unsigned(kk) >> sx) & 1u) : iy;
      jz = sz ? iz * 2 + ((unsigned(kk) >> (sx + sy)) & 1u) : iz;
      cj = j + 1;
    }
    else if (k < 3) {   // BR, TR, BL of transform level il (src/SPECK2D_INT.cpp:150-185)
      jx = k == 2 ? 0u : 1u;
      jy = k == 1 ? 0u : 1u;
      jz = 0;
      cj = il;
    }
    else {
      jx = jy = jz = 0;
      cj = kIFrame | (il - 1);
      child_i = true;
    }
    k++;
    if (s) {
      sg = 1;
      if (!child_i && cj == jC) {   // a size-C set: queue it and skip its bits
        F.qc_node[ntok] = pack(cj, jx, jy, jz);
        F.qc_pos[ntok] = uint16_t(q);
        ntok++;
        q += F.bodyC[q];
      }
      else {   // push (nothing to come back to when this was the last child)
        if (k < nch) {
          F.wk_node[depth] = pack(j, ix, iy, iz);
          F.wk_k[depth] = (unsigned char)k;
          F.wk_sig[depth] = 1;
          depth++;
        }
        j = cj;
        ix = jx; iy = jy; iz = jz;
        k = 0;
        sg = 0;
        nch = nch_of(j);
      }
    }
    else if (child_i)
      F.iset = il - 1;   // stays outside the lists, tested again in the next plane
    else {
      const int cl = lis_of(cj);
      const unsigned slot = F.cnt[cl];
      if (F.off[cl] + slot >= F.off[cl + 1]) {
        F.err |= 1u;
        break;
      }
      gptr(d.lis)[F.off[cl] + slot] = pack(cj, jx, jy, jz);
      F.cnt[cl] = slot + 1;
    }
  }
  if (depth >= 0) {
    F.wk_node[depth] = pack(j, ix, iy, iz);
    F.wk_k[depth] = (unsigned char)k;
    F.wk_sig[depth] = (unsigned char)sg;
  }
  F.wk_depth = depth;
  F.wk_i = i;
  F.wk_q = q;
  F.nqc = ntok;
  S.pos = F.win_base + q;
}

// The walker of a 1D outlier array (SPECK1D_INT_DEC, /root/reference/src/SPECK1D_INT_DEC.cpp:12-125):
// same contract as f_walk<1>, written for the shape of that stream -- with a PWE bound most
// correctors are 1 or 2 quanta, so nearly the whole stream is ONE depth-first descent from the two
// halves of the array in the last plane, a strictly sequential chain of some 10^5 bits that no
// amount of speculation shortens. What can be cut is the cost per bit. In a binary tree the stack is
// implicit: the parent of node ix at depth j is ix >> 1, the child being coded is ix & 1, and a
// frame one returns to has seen a significant child (one only descends into significant sets). So
// the state is (j, ix, k, sg) in registers plus the depth j0 of the staged root; no frame is ever
// parked in shared memory, and the stream is read 32 bits at a time into a register.
#ifndef SPERR_WALK1D_FAST
#define SPERR_WALK1D_FAST 1
#endif
static __device__ void f_walk1d(DecChunk& d, DecShared& S, FastSmem& F)
{
  const unsigned boff = F.boff;
  unsigned cw = 0, cbase = 1u << 30;   // 32 stream bits starting at window position cbase (none yet)
  auto load = [&](unsigned qq) {
    const unsigned a = qq + boff;
    cw = __funnelshift_r(F.bits[a >> 5], F.bits[(a >> 5) + 1], a & 31);
    cbase = qq;
  };
  unsigned q = F.wk_q;
  unsigned ntok = 0;
  int depth = F.wk_depth;   // < 0: between roots
  unsigned i = F.wk_i;
  const unsigned cnt = F.wk_cnt;
  const unsigned i_end = F.rs_first + F.rs_cnt;
  const int jC = F.J - 3;
  int j = 0, k = 0, sg = 0, j0 = 0;
  unsigned ix = 0;
  if (depth >= 0) {   // resume the expansion a previous round left
    j = int(F.wk_node[0] >> 32);
    ix = unsigned(F.wk_node[0]);
    j0 = int(F.wk_node[1]);
    k = F.wk_k[0];
    sg = F.wk_sig[0];
  }
  // sets that stay insignificant are appended to the list of their depth -- not here: a counter
  // look-up, a bounds check and two stores per set are most of what a step of this loop would
  // cost. They are logged (depth, index) in shared memory, at most one per stream bit of the round,
  // and the whole CTA files them afterwards (f_flush_appends); nobody reads those lists before the
  // next plane.
  uint32_t* const alog = reinterpret_cast<uint32_t*>(&F.nxt[0][0]);
  unsigned napp = 0;
  for (;;) {
    if (depth < 0) {
      if (i == cnt)
        break;
      if (q + 1 >= unsigned(kFW) || i >= i_end)
        break;
      const unsigned room = min(min(32u, i_end - i), unsigned(kFW) - 1u - q);
      load(q);
      unsigned u = cw;
      if (room < 32)
        u &= (1u << room) - 1u;
      const unsigned run = u ? unsigned(__ffs(u) - 1) : room;
      i += run;
      q += run;
      if (run == room)
        continue;
      const unsigned r = i - F.rs_first;
      F.rs_gone[r >> 5] |= 1u << (r & 31);
      const node_t nd = F.rs_node[r];
      i++;
      q++;
      depth = 0;
      j = j0 = int(nd >> 32);
      ix = unsigned(nd);
      k = 0;
      sg = 0;
      continue;
    }
    if (k == 2) {   // both halves done: back to the parent, whose first or second half this was
      if (j == j0) {
        depth = -1;
        continue;
      }
      k = int(ix & 1u) + 1;
      ix >>= 1;
      j--;
      sg = 1;
      continue;
    }
    if (q + 1 >= unsigned(kFW) || ntok >= unsigned(kFTok))
      break;
#if SPERR_WALK1D_FAST
    // A freshly entered set with exactly ONE significant size-C set below it -- the usual case in the
    // lower levels of a sparse outlier array -- is finished in one step. On the way down every level
    // costs one bit d (1: the first half is significant, take it; 0: it is not, the second half is
    // significant by inference); after the body of the size-C set, every level that went into its
    // first half tests its second half, one bit each. If those bits are all 0 the h direction bits
    // give the path, every level leaves exactly one insignificant sibling behind (the log is sorted
    // by depth afterwards, stably: entries of different depths may be logged in any order), and the
    // walker is back at this set with both halves done. Otherwise nothing is committed and the set
    // is taken bit by bit as before.
    if (k == 0 && sg == 0) {
      const int h = jC - j;
      if (h >= 2 && q + unsigned(h) < unsigned(kFW)) {
        load(q);
        const unsigned D = cw & ((1u << h) - 1u);   // h <= 24 - 3
        const unsigned qb = q + unsigned(h);        // body of the size-C set at the end of the path
        const unsigned q2 = qb + F.bodyC[qb];
        const unsigned nleft = unsigned(__popc(D));
        if (q2 + nleft < unsigned(kFW)) {
          unsigned rest = 0;
          if (nleft) {
            const unsigned a = q2 + boff;
            rest = __funnelshift_r(F.bits[a >> 5], F.bits[(a >> 5) + 1], a & 31) & ((1u << nleft) - 1u);
          }
          if (rest == 0) {
            const unsigned path = (ix << h) | (__brev(~D) >> (32 - h));   // index of the size-C set
            F.qc_node[ntok] = ((unsigned long long)jC << 32) | path;
            F.qc_pos[ntok] = uint16_t(qb);
            ntok++;
            for (int i = 1; i <= h; i++)
              alog[napp++] = (unsigned(j + i) << 27) | ((path >> (h - i)) ^ 1u);
            q = q2 + nleft;
            k = 2;
            sg = 1;
#if defined(SPERR_EMUL) && defined(SPERR_WALK_COUNT)
            if (threadIdx.x == 0) {
              static unsigned long long fast_levels = 0, fast_hits = 0;
              fast_levels += h;
              ++fast_hits; if ((fast_hits & (fast_hits - 1)) == 0)
                std::fprintf(stderr, "walk1d fast path: %llu hits, %llu levels\n", fast_hits, fast_levels);
            }
#endif
            continue;
          }
        }
      }
    }
#endif
    unsigned s = 1u;
    if (sg != 0 || k == 0) {   // the second half is inferred when the first was insignificant
      if (q - cbase >= 32u)
        load(q);
      s = (cw >> (q - cbase)) & 1u;
      q++;
    }
    const unsigned cix = ix * 2u + unsigned(k);
    const int cj = j + 1;
    k++;
    if (s) {
      sg = 1;
      if (cj == jC) {   // a size-C set: queue it and skip its bits
        F.qc_node[ntok] = ((unsigned long long)cj << 32) | cix;
        F.qc_pos[ntok] = uint16_t(q);
        ntok++;
        q += F.bodyC[q];
      }
      else {
        j = cj;
        ix = cix;
        k = 0;
        sg = 0;
      }
    }
    else
      alog[napp++] = (unsigned(cj) << 27) | cix;   // one store; f_flush_appends puts it into its list
  }
  F.napp = napp;
  if (depth >= 0) {
    F.wk_node[0] = ((unsigned long long)j << 32) | ix;
    F.wk_node[1] = (unsigned long long)j0;
    F.wk_k[0] = (unsigned char)k;
    F.wk_sig[0] = (unsigned char)sg;
  }
  F.wk_depth = depth;
  F.wk_i = i;
  F.wk_q = q;
  F.nqc = ntok;
  S.pos = F.win_base + q;
}

// Survivors of the staged roots [rs_first, wk_i) keep their order: list[wk_w ...] (whole CTA).
// Files the sets the 1D walker logged in this round: entry e = (depth << 27 | index), in log order,
// goes to list `depth` behind the entries of the same depth logged before it (a stable counting
// sort, 1024 entries at a time: ranks inside a warp by match / popc, across warps by a small table).
static __device__ void f_flush_appends(DecChunk& d, FastSmem& F)
{
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned n = F.napp;
  if (n == 0)
    return;
  const uint32_t* const alog = reinterpret_cast<const uint32_t*>(&F.nxt[0][0]);
  uint16_t* const wc = reinterpret_cast<uint16_t*>(&F.nxt[2][0]);   // [32 depths][32 warps]: sets of the batch
  uint32_t* const wt = F.mark;                                      // [32 depths]: batch totals
  for (unsigned b0 = 0; b0 < n; b0 += kDecThreads) {
    const unsigned e = b0 + tid;
    const bool have = e < n;
    const uint32_t v = have ? alog[e] : 0u;
    const unsigned lev = have ? (v >> 27) : 31u;   // idle lanes share a bucket nobody files (depths are < 31)
    const unsigned peers = __match_any_sync(0xffffffffu, lev);
    const unsigned before = unsigned(__popc(peers & ((1u << lane) - 1u)));
    for (int i = tid; i < 32 * 32; i += kDecThreads)
      wc[i] = 0;
    block_sync();
    if (have && before == 0)
      wc[lev * 32 + warp] = uint16_t(__popc(peers));
    block_sync();
    {   // warp l scans depth l over the 32 warps: exclusive prefix back into the table, total to wt
      const unsigned c = wc[warp * 32 + lane];
      unsigned inc = c;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
          inc += t;
      }
      wc[warp * 32 + lane] = uint16_t(inc - c);
      if (lane == 31) {
        wt[warp] = inc;
        if (inc && F.off[warp] + F.cnt[warp] + inc > F.off[warp + 1])
          F.err |= 1u;
      }
    }
    block_sync();
    if (have && !F.err)
      gptr(d.lis)[F.off[lev] + F.cnt[lev] + wc[lev * 32 + warp] + before] = ((unsigned long long)lev << 32) | (v & 0x7ffffffu);
    block_sync();
    if (tid < 31)
      F.cnt[tid] += wt[tid];
    block_sync();
  }
}

static __device__ void f_compact_roots(FastSmem& F, node_t* list)
{
  const int tid = threadIdx.x;
#ifdef SPERR_DEC_VOLATILE
  const unsigned n = *(volatile unsigned*)&F.wk_i - *(volatile unsigned*)&F.rs_first;
  const unsigned w0 = *(volatile unsigned*)&F.wk_w;
#else
  const unsigned n = F.wk_i - F.rs_first;   // roots visited in this round
  const unsigned w0 = F.wk_w;
#endif
  unsigned total = 0;
  for (unsigned b0 = 0; b0 < n; b0 += kDecThreads) {
    const unsigned t = b0 + tid;
    const bool keep = t < n && !((F.rs_gone[t >> 5] >> (t & 31)) & 1u);
    const unsigned long long ex = f_block_scan(F, keep ? 1ull : 0ull);
    if (keep && (w0 + total + unsigned(ex) != F.rs_first + t))
      list[w0 + total + unsigned(ex)] = F.rs_node[t];
    total += unsigned(F.scan_total);
  }
  block_sync();
  if (tid == 0)
    F.wk_w = w0 + total;
}

// The LIS part of one bit-plane: the lists of the three smallest set sizes (every CTA of the
// cluster) ...
template <bool CL>
static __device__ void dec_lis_chains(DecChunk& d, DecShared& S, FastSmem& F, int n_plane)
{
  for (int kind = 0; kind < 3; kind++) {
    f_list_by_chain<CL>(d, S, F, kind, n_plane);
    if (F.err)
      return;
  }
}

// ... and the lists of the larger sets (one CTA)
#ifdef SPERR_DEC_FORCEINLINE
__forceinline__
#endif
static __device__ void dec_lis_walk(DecChunk& d, DecShared& S, FastSmem& F, int n_plane)
{
  const int tid = threadIdx.x;
  bool have_window = false;
  // lj == 0 (2D only): the set I, tested after all lists (src/SPECK2D_INT.cpp:54-57), as a
  // one-entry pseudo list
  for (int lj = max(F.J - 4, 0); lj >= (F.is2d ? 0 : 1); lj--) {
    const bool iphase = lj == 0;
    const int lis = f_lis(F, iphase ? 1 : lj);
    const unsigned cnt = iphase ? (F.iset > 0 ? 1u : 0u) : F.cnt[lis];
    if (cnt == 0)
      continue;
    if (tid == 0) {
      F.wk_j = iphase ? 1 : lj;
      F.wk_i = 0;
      F.wk_w = 0;
      F.wk_cnt = cnt;
      F.wk_depth = -1;
    }
    block_sync();
    for (;;) {
      if (!have_window || F.wk_q + 1 >= unsigned(kFW)) {
        f_build_window(d, F, S.pos, 2, false);
        if (tid == 0) {
          F.win_base = S.pos;
          F.wk_q = 0;
        }
        have_window = true;
      }
      // stage the next roots of the list (the walker never reads the list itself)
      const unsigned first = F.wk_i;
      const unsigned nst = min(unsigned(kFRoots), cnt - first);
      block_sync();
      for (unsigned t = tid; t < nst; t += kDecThreads)
        F.rs_node[t] = iphase ? ((unsigned long long)(0x100 | F.iset) << 32)
                              : f_list_load(F, gptr(d.lis) + F.off[lis] + first + t);
      for (unsigned t = tid; t < unsigned(kFRoots / 32); t += kDecThreads)
        F.rs_gone[t] = 0;
      if (tid == 0) {
        F.rs_first = first;
        F.rs_cnt = nst;
      }
      block_sync();
      const long long f_tw = F_CLOCK();
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        if (d.kind == 1)
#ifdef SPERR_OLD_WALK1D
          f_walk<1>(d, S, F);
#else
          f_walk1d(d, S, F);
#endif
        else if (d.kind == 2)
          f_walk<2>(d, S, F);
        else
          f_walk<0>(d, S, F);
      }
      block_sync();
      if (d.kind == 1)
        f_flush_appends(d, F);
      if (!iphase)
        f_compact_roots(F, gptr(d.lis) + F.off[lis]);
      block_sync();
      const long long f_te = F_CLOCK();
      if (tid == 0)
        F.prof[4] += (unsigned long long)(f_te - f_tw);
      f_expand_round<false>(d, S, F, 2, n_plane);
      block_sync();
      if (tid == 0)
        F.prof[3] += (unsigned long long)(F_CLOCK() - f_te);
      if (F.err)
        return;
      if (F.wk_depth < 0 && F.wk_i == cnt)
        break;
    }
    if (DEC_SERIAL(tid) && d.dbgw_n < 16) {
      DEC_SERIAL_CONVERGE();
      unsigned* const r = d.dbgw[d.dbgw_n++];
      r[0] = unsigned(n_plane); r[1] = unsigned(lj); r[2] = cnt; r[3] = F.wk_i; r[4] = F.wk_w;
      r[5] = unsigned(lis); r[6] = F.wk_q; r[7] = F.rs_cnt;
    }
    if (tid == 0 && !iphase)
      F.cnt[lis] = F.wk_w;   // survivors; the sets created in this plane went to deeper lists
    block_sync();
  }
}

// ---- the kernel -----------------------------------------------------------------------------------

// Cluster mode: the leader (rank 0) runs the LIP pass, the walker and the end-of-plane bookkeeping
// alone and publishes the decoder state; the other CTAs pick it up for the chain phases.
static __device__ void f_state_sync(DecChunk& d, DecShared& S, FastSmem& F)
{
  const int tid = threadIdx.x;
  ClusterBox* const box = d.box;
  if (F.rank == 0) {
    if (tid == 0) {
      mb_st(&box->pos, S.pos);
      mb_st(&box->klip, S.klip);
      mb_st(&box->klsp, S.klsp);
      mb_st(&box->knew, S.knew);
      mb_st(&box->iset, F.iset);
      mb_st(&box->go, F.go);
      mb_st(&box->plane, F.plane);
      mb_st(&box->err, F.err);
    }
    for (int l = tid; l < d.nlis; l += kDecThreads)
      mb_st(&box->cnt[l], F.cnt[l]);
  }
  cluster_sync();
  if (F.rank != 0) {
    if (tid == 0) {
      S.pos = mb_ld(&box->pos);
      S.klip = mb_ld(&box->klip);
      S.klsp = mb_ld(&box->klsp);
      S.knew = mb_ld(&box->knew);
      F.iset = mb_ld(&box->iset);
      F.go = mb_ld(&box->go);
      F.plane = mb_ld(&box->plane);
      F.err = mb_ld(&box->err);
    }
    for (int l = tid; l < d.nlis; l += kDecThreads)
      F.cnt[l] = mb_ld(&box->cnt[l]);
  }
  cluster_sync();   // read before the leader publishes again
}

// CL: launched as clusters of chunks[.].R CTAs per job (every job of the launch has the same R).
template <bool CL>
static __global__ void __launch_bounds__(kDecThreads) k_speck_decode_fast(DecChunk* chunks, int R)
{
  __shared__ DecShared S;
  DYN_SMEM(FastSmem, Fp);
  FastSmem& F = *Fp;
  const unsigned c = CL ? blockIdx.x / unsigned(R) : blockIdx.x;
  const int rank = CL ? int(cluster_rank()) : 0;
  if (chunks[c].skip || chunks[c].planes == 0 || !chunks[c].pow2 || chunks[c].R != R)
    return;   // the same for every CTA of a cluster
  const int tid = threadIdx.x;
#ifdef SPERR_DEC_ZERO_SMEM
  for (unsigned i = tid; i < sizeof(FastSmem) / 4; i += blockDim.x)
    reinterpret_cast<unsigned*>(&F)[i] = 0;
  for (unsigned i = tid; i < sizeof(DecShared) / 4; i += blockDim.x)
    reinterpret_cast<unsigned*>(&S)[i] = 0;
  block_sync();
#endif
  // The job descriptor is worked on in shared memory and written back at the end. Read through
  // the global struct, every pointer in it (bits, pl, lip, lis, ...) has to be fetched again after
  // each global store, which may alias it: one more dependent load in front of nearly every access
  // of a kernel that is bound by exactly such chains. Shared memory cannot alias the global arrays.
  __shared__ DecChunk sd;
  static_assert(sizeof(DecChunk) % 8 == 0, "DecChunk is copied in 8-byte words");
  for (unsigned i = tid; i < sizeof(DecChunk) / 8; i += blockDim.x)
    reinterpret_cast<unsigned long long*>(&sd)[i] = reinterpret_cast<const unsigned long long*>(&chunks[c])[i];
  block_sync();
  DecChunk& d = sd;
  if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
    S.pos = 0;
    S.klip = 0;
    S.klsp = 0;
    S.knew = 0;
    S.endpos = 0;
    F.Dx = d.Dx; F.Dy = d.Dy; F.Dz = d.Dz;
    F.J = max(d.Dx, max(d.Dy, d.Dz));
    F.nx = d.nx; F.ny = d.ny;
    F.is2d = d.kind == 2;
    F.iset = d.kind == 2 ? int(d.iset) : 0;
    F.err = 0;
    F.R = R;
    F.rank = rank;
    F.go = 1;
    F.plane = d.planes - 1;
    F.nqa = F.nqb = F.nqc = 0;
    for (int k = 0; k < 8; k++)
      F.prof[k] = 0;
    for (int l = 0; l <= d.nlis; l++)
      F.off[l] = gptr(d.lis_off)[l];
    for (int l = 0; l < d.nlis; l++)
      F.cnt[l] = 0;
  }
  f_build_luts(F);
  block_sync();
  if (DEC_SERIAL(tid) && rank == 0) {
      DEC_SERIAL_CONVERGE();
    for (int r = 0; r < d.nroots; r++) {
      const unsigned long long nd = d.roots[r];
      const int lis = f_lis(F, int(nd >> 32));
      gptr(d.lis)[F.off[lis] + F.cnt[lis]] = nd;
      F.cnt[lis]++;
    }
  }
  block_sync();
  for (;;) {
    if (CL)
      f_state_sync(d, S, F);
    block_sync();
    if (!F.go)
      break;
    const int n = F.plane;
    {   // LIP part: the leader tokenises the bit string, every CTA matches its share of the mask
      F_TIC();
#ifdef SPERR_CL_LIP_LEADER
      if (rank == 0)
        dec_lip_pass(d, S, n);
      if (CL)
        f_state_sync(d, S, F);
#else
      dec_lip_pass(d, S, n, CL ? R : 1, rank, CL ? &d.box->lip : nullptr);
#endif
      block_sync();
      F_TOC(F, 0);
    }
    auto trace = [&](int stage) {   // debugging aid, leader only
      if (DEC_SERIAL(tid) && rank == 0 && n < kMaxPlanes) {
      DEC_SERIAL_CONVERGE();
        unsigned long long sets = 0;
        for (int l = 0; l < d.nlis; l++)
          sets += F.cnt[l];
        d.dbg[n][stage][0] = S.pos;
        d.dbg[n][stage][1] = (S.klip << 40) ^ (S.knew << 20) ^ sets;
        const int pi = d.planes - 1 - n;
        if (pi < 2)
          for (int l = 0; l < d.nlis && l < 32; l++)
            d.dbgl[pi][stage][l] = F.cnt[l];
      }
    };
    trace(0);
    dec_lis_chains<CL>(d, S, F, n);
    block_sync();
    trace(1);
    if (rank == 0) {
      if (!F.err)
        dec_lis_walk(d, S, F, n);
      block_sync();
      trace(2);
      if (DEC_SERIAL(tid)) {
      DEC_SERIAL_CONVERGE();
        const bool more = !F.err && dec_plane_end(d, S, n);
        F.go = (more && n > 0) ? 1 : 0;
        F.plane = n - 1;
      }
      block_sync();
    }
  }
  if (rank != 0)
    return;
  if (tid == 0) {
    d.err = F.err;
    for (int k = 0; k < 8; k++)
      d.prof[k] = F.prof[k];
  }
  block_sync();
  for (unsigned i = tid; i < sizeof(DecChunk) / 8; i += blockDim.x)
    reinterpret_cast<unsigned long long*>(&chunks[c])[i] = reinterpret_cast<const unsigned long long*>(&sd)[i];
}

}  // namespace sperr_b200
