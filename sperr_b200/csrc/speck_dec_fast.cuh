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
constexpr int kFTok = 1024;           // tokens the walker may queue per round (one per thread)
constexpr int kFMaxDepth = 32;

struct FastSmem {
  uint32_t bits[kFBits / 32 + 4];     // window of the stream, word aligned; bit q is at q + boff
  uint16_t bodyC[kFW + 8];
  uint8_t bodyB[kFBodyB + 8];
  uint8_t bodyA[kFBodyA + 8];
  uint16_t nxt[2][kFNext + 8];
  uint32_t mark[kFW / 32];
  uint16_t T1[256], T2[2][256];       // pixel-set tables: sig(4) | sign(4) << 4 | bits << 8
  unsigned long long tok_node[kFTok];
  uint32_t tok_pos[kFTok];
  unsigned long long off[kMaxLis + 1];
  unsigned cnt[kMaxLis];
  unsigned long long wsum[kDecWarps];
  unsigned long long scan_total;
  unsigned boff;
  unsigned cutpos;
  // geometry
  int Dx, Dy, Dz, J;
  unsigned nx, ny;
  // walker (thread 0) state, kept here so that it can be suspended between windows
  int wk_j;                 // depth of the list being visited
  unsigned wk_i, wk_w, wk_cnt;
  int wk_depth;             // -1: between roots
  unsigned long long wk_node[kFMaxDepth];
  unsigned char wk_k[kFMaxDepth], wk_sig[kFMaxDepth];
  int wk_done;
  unsigned ntok;
  unsigned err;
  unsigned long long prof[8];
};

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

// ---- geometry of power-of-two trees -----------------------------------------------------------
// A set at depth j is the box (ix, iy, iz) of the grid with 2^min(j, D_a) cells along axis a; it is
// stored as (j << 32 | linear index), its list index is sum_a min(j, D_a).

__device__ __forceinline__ int f_lis(const FastSmem& F, int j)
{
  return min(j, F.Dx) + min(j, F.Dy) + min(j, F.Dz);
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
  __syncthreads();   // previous users of wsum / scan_total are done
  if (lane == 31)
    F.wsum[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    const unsigned long long w = F.wsum[lane];
    const unsigned long long winc = warp_incl_scan(w, lane);
    F.wsum[lane] = winc - w;
    if (lane == 31)
      F.scan_total = winc;
  }
  __syncthreads();
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
                                           int kinds);
static __device__ void f_build_window(const DecChunk& d, FastSmem& F, unsigned long long base, int kinds)
{
  F_TIC();
  f_build_window_impl(d, F, base, kinds);
  __syncthreads();
  F_TOC(F, 1);
  if (threadIdx.x == 0)
    F.prof[6]++;
}
static __device__ void f_build_window_impl(const DecChunk& d, FastSmem& F, unsigned long long base,
                                           int kinds)
{
  const int tid = threadIdx.x;
  __syncthreads();   // every reader of the previous window is done
  const unsigned long long w0 = base >> 5;
  const int nbits = kinds == 0 ? kFW + 64 : (kinds == 1 ? kFW + 256 : kFBits);
  for (int i = tid; i < nbits / 32 + 3; i += kDecThreads) {
    const unsigned long long w = w0 + i;
    F.bits[i] = w < d.stage_words ? d.bits[w] : 0u;
  }
  if (tid == 0)
    F.boff = unsigned(base & 31);
  __syncthreads();
  const int J = F.J;
  const int nchA = f_nch(F, J - 1);
  const int nA = kinds == 0 ? kFW + 8 : (kinds == 1 ? kFW + 144 : kFBodyA);
  for (int q = tid; q < nA; q += kDecThreads) {
    unsigned sm, gm;
    F.bodyA[q] = uint8_t(f_pixels(F, q, nchA, sm, gm));
  }
  if (kinds == 0)
    return;
  __syncthreads();
  const int nchB = f_nch(F, J - 2);
  const int nB = kinds == 1 ? kFW + 8 : kFBodyB;
  for (int q = tid; q < nB; q += kDecThreads) {
    unsigned pos = q;
    int c = 0;
    for (int k = 0; k < nchB; k++) {
      const bool need = c != 0 || k != nchB - 1;
      const unsigned s = need ? f_bit(F, pos++) : 1u;
      if (s) {
        c = 1;
        pos += F.bodyA[pos];
      }
    }
    F.bodyB[q] = uint8_t(pos - q);
  }
  if (kinds == 1)
    return;
  __syncthreads();
  const int nchC = f_nch(F, J - 3);
  for (int q = tid; q < kFW + 8; q += kDecThreads) {
    unsigned pos = q;
    int c = 0;
    for (int k = 0; k < nchC; k++) {
      const bool need = c != 0 || k != nchC - 1;
      const unsigned s = need ? f_bit(F, pos++) : 1u;
      if (s) {
        c = 1;
        pos += F.bodyB[pos];
      }
    }
    F.bodyC[q] = uint16_t(pos - q);
  }
}

// ---- token expansion (one thread per token) -----------------------------------------------------

struct FastExp {
  unsigned q;              // cursor (window position)
  unsigned nA, nB, nlip;   // sets left for the two deepest lists, pixels that entered the LIP
  unsigned long long* scr; // this thread's staging: [0, 16) list B, [16, 96) list A
};

// pixels of a significant bottom-level set (depth J - 1)
static __device__ void f_expand_pixels(const DecChunk& d, const FastSmem& F, FastExp& e, unsigned ix,
                                       unsigned iy, unsigned iz)
{
  const int j = F.J - 1;
  const int sx = j < F.Dx, sy = j < F.Dy, sz = j < F.Dz;
  const int nch = 1 << (sx + sy + sz);
  unsigned sigm, sgnm;
  e.q += f_pixels(F, e.q, nch, sigm, sgnm);
  e.nlip += unsigned(nch - __popc(sigm));
  const unsigned x0 = sx ? ix * 2 : ix, y0 = sy ? iy * 2 : iy, z0 = sz ? iz * 2 : iz;
  const unsigned long long nx = F.nx, nxy = (unsigned long long)F.nx * F.ny;
  const unsigned long long base = (unsigned long long)z0 * nxy + (unsigned long long)y0 * nx + x0;
  // pixels that differ in x only are neighbours in the masks: one update per row
  const int per_row = sx ? 2 : 1;
  const int rows = nch / per_row;
  for (int r = 0; r < rows; r++) {
    const unsigned cy = sy ? (unsigned(r) & 1u) : 0u;
    const unsigned cz = sz ? ((unsigned(r) >> sy) & 1u) : 0u;
    const unsigned long long i = base + cy * nx + cz * nxy;
    const unsigned rm = sx ? 3u : 1u;
    const unsigned s = (sigm >> (r * per_row)) & rm, g = (sgnm >> (r * per_row)) & rm;
    const unsigned sh = unsigned(i & 31);   // x0 is even when the row has two pixels: same word
    const unsigned long long w = i >> 5;
    if (s)
      atomicOr(&d.newm[w], s << sh);
    if (s & ~g)
      atomicAnd(&d.signs[w], ~((s & ~g) << sh));
    if (rm & ~s)
      atomicOr(&d.lip[w], (rm & ~s) << sh);
  }
}

// KIND 0: depth J-1 (children are pixels), 1: depth J-2, 2: depth J-3
template <int KIND>
static __device__ void f_expand(const DecChunk& d, const FastSmem& F, FastExp& e, unsigned ix,
                                unsigned iy, unsigned iz)
{
  if (KIND == 0) {
    f_expand_pixels(d, F, e, ix, iy, iz);
    return;
  }
  const int j = F.J - 1 - KIND;
  const int nch = f_nch(F, j);
  int c = 0;
  for (int k = 0; k < nch; k++) {
    const bool need = c != 0 || k != nch - 1;
    const unsigned s = need ? f_bit(F, e.q++) : 1u;
    unsigned jx, jy, jz;
    f_child(F, j, ix, iy, iz, k, jx, jy, jz);
    if (s) {
      c = 1;
      f_expand<(KIND > 0 ? KIND - 1 : 0)>(d, F, e, jx, jy, jz);
    }
    else if (KIND == 1)
      e.scr[16 + e.nA++] = f_pack(F, j + 1, jx, jy, jz);
    else
      e.scr[e.nB++] = f_pack(F, j + 1, jx, jy, jz);
  }
}

// Appends what the expanders staged to the two deepest lists, in thread (= token) order.
static __device__ void f_commit(DecChunk& d, DecShared& S, FastSmem& F, const FastExp& e)
{
  const unsigned long long packed =
      (unsigned long long)e.nA | ((unsigned long long)e.nB << 21) | ((unsigned long long)e.nlip << 42);
  const unsigned long long ex = f_block_scan(F, packed);
  const unsigned long long tot = F.scan_total;
  const unsigned exA = unsigned(ex) & 0x1fffffu, exB = unsigned(ex >> 21) & 0x1fffffu;
  const unsigned totA = unsigned(tot) & 0x1fffffu, totB = unsigned(tot >> 21) & 0x1fffffu;
  const int lisA = f_lis(F, F.J - 1), lisB = f_lis(F, F.J - 2);
  const unsigned long long pa = F.off[lisA] + F.cnt[lisA], pb = F.off[lisB] + F.cnt[lisB];
  const bool okA = pa + totA <= F.off[lisA + 1], okB = pb + totB <= F.off[lisB + 1];
  if (okA)
    for (unsigned i = 0; i < e.nA; i++)
      d.lis[pa + exA + i] = e.scr[16 + i];
  if (okB)
    for (unsigned i = 0; i < e.nB; i++)
      d.lis[pb + exB + i] = e.scr[i];
  __syncthreads();
  if (threadIdx.x == 0) {
    if (!okA || !okB)
      F.err |= 1u;
    F.cnt[lisA] += totA;
    F.cnt[lisB] += totB;
    S.klip += tot >> 42;
  }
  __syncthreads();
}

// ---- phase A: the lists of the three smallest set sizes, by token chains -------------------------

static __device__ void f_list_by_chain(DecChunk& d, DecShared& S, FastSmem& F, int kind)
{
  const int tid = threadIdx.x;
  const int j = F.J - 1 - kind;
  const int lis = f_lis(F, j);
  const unsigned m = F.cnt[lis];
  if (m == 0)
    return;
  node_t* const list = d.lis + F.off[lis];
  unsigned i0 = 0, wsurv = 0;
  while (i0 < m) {
    f_build_window(d, F, S.pos, kind);
    const long long f_tc = F_CLOCK();
    // token length of every window position; positions beyond the window absorb
    for (int q = tid; q < kFNext; q += kDecThreads) {
      unsigned nx = q;
      if (q < kFW) {
        unsigned len = 1;
        if (f_bit(F, q))
          len += kind == 0 ? F.bodyA[q + 1] : (kind == 1 ? F.bodyB[q + 1] : F.bodyC[q + 1]);
        nx = q + len;
      }
      F.nxt[0][q] = uint16_t(nx);
    }
    for (int i = tid; i < kFW / 32; i += kDecThreads)
      F.mark[i] = i == 0 ? 1u : 0u;
    __syncthreads();
    // pointer doubling: after round r the marks cover the first 2^(r+1) tokens of the chain
    int cur = 0;
    for (int r = 0; r < 14; r++) {
      const uint16_t* cn = F.nxt[cur];
      uint16_t* nn = F.nxt[cur ^ 1];
      const unsigned mb = (F.mark[tid >> 2] >> ((tid & 3) * 8)) & 0xffu;   // my 8 positions
      unsigned t = mb;
      while (t) {
        const int b = __ffs(t) - 1;
        t &= t - 1;
        const unsigned tgt = cn[tid * 8 + b];
        if (tgt < kFW)
          atomicOr(&F.mark[tgt >> 5], 1u << (tgt & 31));
      }
      for (int q = tid; q < kFNext; q += kDecThreads)
        nn[q] = q < kFW ? cn[cn[q]] : uint16_t(q);
      __syncthreads();
      cur ^= 1;
      if (F.nxt[cur][0] >= kFW)   // 2^(r+1) steps leave the window: every token start is marked
        break;
    }
    const unsigned exitpos = F.nxt[cur][0];
    // rank of my marks = index of the root they belong to
    const unsigned mb = (F.mark[tid >> 2] >> ((tid & 3) * 8)) & 0xffu;
    unsigned long long ex = f_block_scan(F, (unsigned long long)__popc(mb));
    const unsigned total_marks = unsigned(F.scan_total);
    const unsigned T = min(total_marks, m - i0);
    // collect my tokens: survivors keep their place in the list, significant ones are expanded
    node_t surv[8], sigs[8];
    unsigned sigq[8];
    int ns = 0, ng = 0;
    {
      unsigned t = mb, r = unsigned(ex);
      while (t) {
        const int b = __ffs(t) - 1;
        t &= t - 1;
        const unsigned q = tid * 8 + b;
        if (r < T) {
          const node_t nd = list[i0 + r];
          if (f_bit(F, q)) {
            sigs[ng] = nd;
            sigq[ng++] = q + 1;
          }
          else
            surv[ns++] = nd;
        }
        else if (r == T)
          F.cutpos = q;   // first position after this list's last token of the window
        r++;
      }
    }
    ex = f_block_scan(F, (unsigned long long)ns);   // barrier inside: every root has been read
    const long long f_te = F_CLOCK();
    if (tid == 0)
      F.prof[2] += (unsigned long long)(f_te - f_tc);
    const unsigned tot_surv = unsigned(F.scan_total);
    for (int s = 0; s < ns; s++)
      list[wsurv + unsigned(ex) + s] = surv[s];
    FastExp e;
    e.nA = e.nB = e.nlip = 0;
    e.scr = d.scr + (size_t)tid * kFastScrPerThread;
    for (int g = 0; g < ng; g++) {
      int jj;
      unsigned ix, iy, iz;
      f_unpack(F, sigs[g], jj, ix, iy, iz);
      e.q = sigq[g];
      if (kind == 0)
        f_expand<0>(d, F, e, ix, iy, iz);
      else if (kind == 1)
        f_expand<1>(d, F, e, ix, iy, iz);
      else
        f_expand<2>(d, F, e, ix, iy, iz);
    }
    f_commit(d, S, F, e);
    if (tid == 0) {
      F.prof[3] += (unsigned long long)(F_CLOCK() - f_te);
      S.pos += T == total_marks ? exitpos : F.cutpos;
    }
    i0 += T;
    wsurv += tot_surv;
    __syncthreads();
    if (F.err)
      return;
  }
  if (tid == 0)
    F.cnt[lis] = wsurv;
  __syncthreads();
}

// ---- phase B: larger sets, one thread walks the top of the tree ----------------------------------

// Runs until the lists are exhausted (wk_done) or the window / token queue is used up.
static __device__ void f_walk(DecChunk& d, DecShared& S, FastSmem& F)
{
  const unsigned long long base = S.pos;
  unsigned q = 0;            // window position
  unsigned ntok = 0;
  int depth = F.wk_depth;
  int lj = F.wk_j;
  unsigned i = F.wk_i, w = F.wk_w, cnt = F.wk_cnt;
  const int jC = F.J - 3;
  for (;;) {
    if (depth < 0) {
      if (i == cnt) {   // list finished: survivors are compacted, new sets of this plane follow them
        if (lj >= 1)
          F.cnt[f_lis(F, lj)] = w;
        lj--;
        if (lj < 1) {
          F.wk_done = 1;
          break;
        }
        i = 0;
        w = 0;
        cnt = F.cnt[f_lis(F, lj)];
        continue;
      }
      if (q + 1 >= unsigned(kFW))
        break;
      node_t* const list = d.lis + F.off[f_lis(F, lj)];
      const node_t nd = list[i++];
      if (f_bit(F, q++) == 0) {
        list[w++] = nd;
        continue;
      }
      depth = 0;
      F.wk_node[0] = nd;
      F.wk_k[0] = 0;
      F.wk_sig[0] = 0;
      continue;
    }
    int j;
    unsigned ix, iy, iz;
    f_unpack(F, F.wk_node[depth], j, ix, iy, iz);
    const int nch = f_nch(F, j);
    const int k = F.wk_k[depth];
    if (k == nch) {
      depth--;
      continue;
    }
    if (q + 1 >= unsigned(kFW) || ntok >= unsigned(kFTok))
      break;
    F.wk_k[depth] = (unsigned char)(k + 1);
    const bool need = F.wk_sig[depth] != 0 || k != nch - 1;
    const unsigned s = need ? f_bit(F, q++) : 1u;
    unsigned jx, jy, jz;
    f_child(F, j, ix, iy, iz, k, jx, jy, jz);
    const node_t child = f_pack(F, j + 1, jx, jy, jz);
    if (s) {
      F.wk_sig[depth] = 1;
      if (j + 1 == jC) {   // a bodyC-sized set: queue it and skip its bits
        F.tok_node[ntok] = child;
        F.tok_pos[ntok] = q;
        ntok++;
        q += F.bodyC[q];
      }
      else {
        depth++;
        F.wk_node[depth] = child;
        F.wk_k[depth] = 0;
        F.wk_sig[depth] = 0;
      }
    }
    else {
      const int cl = f_lis(F, j + 1);
      const unsigned slot = F.cnt[cl];
      if (F.off[cl] + slot >= F.off[cl + 1]) {
        F.err |= 1u;
        break;
      }
      d.lis[F.off[cl] + slot] = child;
      F.cnt[cl] = slot + 1;
    }
  }
  F.wk_depth = depth;
  F.wk_j = lj;
  F.wk_i = i;
  F.wk_w = w;
  F.wk_cnt = cnt;
  F.ntok = ntok;
  S.pos = base + q;
}

// The LIS part of one bit-plane.
static __device__ void dec_lis_fast(DecChunk& d, DecShared& S, FastSmem& F)
{
  const int tid = threadIdx.x;
  for (int kind = 0; kind < 3; kind++) {
    f_list_by_chain(d, S, F, kind);
    if (F.err)
      return;
  }
  if (F.J - 4 < 1)
    return;
  if (tid == 0) {
    F.wk_j = F.J - 4;
    F.wk_i = 0;
    F.wk_w = 0;
    F.wk_cnt = F.cnt[f_lis(F, F.J - 4)];
    F.wk_depth = -1;
    F.wk_done = 0;
  }
  __syncthreads();
  for (;;) {
    f_build_window(d, F, S.pos, 2);
    const long long f_tw = F_CLOCK();
    if (tid == 0)
      f_walk(d, S, F);
    __syncthreads();
    const long long f_te = F_CLOCK();
    if (tid == 0)
      F.prof[4] += (unsigned long long)(f_te - f_tw);
    FastExp e;
    e.nA = e.nB = e.nlip = 0;
    e.scr = d.scr + (size_t)tid * kFastScrPerThread;
    if (unsigned(tid) < F.ntok) {
      int jj;
      unsigned ix, iy, iz;
      f_unpack(F, F.tok_node[tid], jj, ix, iy, iz);
      e.q = F.tok_pos[tid];
      f_expand<2>(d, F, e, ix, iy, iz);
    }
    f_commit(d, S, F, e);
    if (tid == 0)
      F.prof[3] += (unsigned long long)(F_CLOCK() - f_te);
    if (F.err || F.wk_done)
      break;
  }
}

// ---- the kernel -----------------------------------------------------------------------------------

static __global__ void __launch_bounds__(kDecThreads) k_speck_decode_fast(DecChunk* chunks)
{
  __shared__ DecShared S;
  DYN_SMEM(FastSmem, Fp);
  FastSmem& F = *Fp;
  const unsigned c = blockIdx.x;
  DecChunk& d = chunks[c];
  if (d.skip || d.planes == 0 || !d.pow2)
    return;
  const int tid = threadIdx.x;
  if (tid == 0) {
    S.pos = 0;
    S.klip = 0;
    S.klsp = 0;
    S.endpos = 0;
    F.Dx = d.Dx; F.Dy = d.Dy; F.Dz = d.Dz;
    F.J = max(d.Dx, max(d.Dy, d.Dz));
    F.nx = d.nx; F.ny = d.ny;
    F.err = 0;
    for (int k = 0; k < 8; k++)
      F.prof[k] = 0;
    for (int l = 0; l <= d.nlis; l++)
      F.off[l] = d.lis_off[l];
    for (int l = 0; l < d.nlis; l++)
      F.cnt[l] = 0;
  }
  f_build_luts(F);
  __syncthreads();
  if (tid == 0) {
    for (int r = 0; r < d.nroots; r++) {
      const unsigned long long nd = d.roots[r];
      const int lis = f_lis(F, int(nd >> 32));
      d.lis[F.off[lis] + F.cnt[lis]] = nd;
      F.cnt[lis]++;
    }
  }
  __syncthreads();
  int n = d.planes - 1;
  bool pending_new = false;
  for (int bp = 0; bp < d.planes; bp++, n--) {
    {
      F_TIC();
      dec_lip_pass(d, S);
      __syncthreads();
      F_TOC(F, 0);
    }
    dec_lis_fast(d, S, F);
    __syncthreads();
    if (F.err) {
      if (tid == 0)
        d.err = F.err;
      return;
    }
    if (S.pos >= d.avail) {
      pending_new = true;
      break;
    }
    F_TIC();
    dec_refine_pass(d, S, n, true);
    F_TOC(F, 5);
    if (S.pos >= d.avail)
      break;
  }
  if (pending_new)
    dec_refine_pass(d, S, n, false);
  if (tid == 0)
    for (int k = 0; k < 8; k++)
      d.prof[k] = F.prof[k];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------

// out[w] = little-endian word w of the byte string src[0, len), zero beyond it
static __global__ void k_stage_bits(const DecChunk* chunks, const unsigned char* const* srcs,
                                    const unsigned long long* lens, const unsigned long long* words)
{
  const unsigned c = blockIdx.y;
  const unsigned char* src = srcs[c];
  const unsigned long long len = lens[c], nw = words[c];
  uint32_t* out = const_cast<uint32_t*>(chunks[c].bits);
  const unsigned long long avail = chunks[c].avail;
  for (unsigned long long w = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; w < nw;
       w += (unsigned long long)gridDim.x * blockDim.x) {
    uint32_t v = 0;
    for (int b = 0; b < 4; b++) {
      const unsigned long long i = w * 4 + b;
      if (i < len)
        v |= uint32_t(src[i]) << (8 * b);
    }
    if (w * 32 >= avail)   // bits the stream does not really hold read as zero
      v = 0;
    else if (avail - w * 32 < 32)
      v &= (1u << (avail - w * 32)) - 1u;
    out[w] = v;
  }
}

// Decodes every job; on return w.h[c].lsp is the final significance mask of job c.
template <class T>
void run_decoder(DecWork& w, const std::vector<DecJob>& jobs, const typename T::Data& tree,
                 cudaStream_t st)
{
  const int nj = int(jobs.size());
  if (nj == 0)
    return;
  size_t mask_words = 0, lis_entries = 0, cnt_entries = 0, stage_words = 0;
  bool any_fast = false, any_slow = false;
  std::vector<size_t> mw(nj), sw(nj);
  for (int c = 0; c < nj; c++) {
    const DecJob& j = jobs[c];
    mw[c] = j.skip ? 0 : (size_t(j.n + 31) / 32 + 4);
    sw[c] = j.skip ? 0 : (size_t(j.payload_bytes) / 4 + 4);
    mask_words += 5 * mw[c];
    lis_entries += j.skip ? 0 : j.lis_total;
    cnt_entries += j.skip ? 0 : size_t(j.nlis + 1);
    stage_words += sw[c];
    if (!j.skip)
      (j.pow2 ? any_fast : any_slow) = true;
  }
  if (any_fast)
    w.scr.reserve(size_t(nj) * kDecThreads * kFastScrPerThread * 8);
  w.masks.reserve(mask_words * 4 + 16);
  w.lis.reserve(lis_entries * 8 + 16);
  w.lis_cnt.reserve(cnt_entries * 4 + 16);
  w.stage.reserve(stage_words * 4 + 16);
  rt::dset(w.masks.p, 0, mask_words * 4, st);
  w.h.assign(nj, DecChunk());
  std::vector<const unsigned char*> srcs(nj);
  std::vector<unsigned long long> lens(nj), words(nj);
  size_t om = 0, ol = 0, oc = 0, os = 0, max_words = 1;
  for (int c = 0; c < nj; c++) {
    const DecJob& j = jobs[c];
    DecChunk& d = w.h[c];
    std::memset(&d, 0, sizeof(d));
    d.skip = j.skip ? 1 : 0;
    srcs[c] = j.d_payload;
    lens[c] = j.skip ? 0 : j.payload_bytes;
    words[c] = sw[c];
    if (j.skip)
      continue;
    d.n = j.n;
    d.shape = j.shape;
    d.planes = j.planes;
    d.avail = std::min<unsigned long long>(j.total_bits, j.payload_bytes * 8ull);
    d.wide = j.wide;
    d.mag = j.mag;
    d.signs = j.signs;
    uint32_t* m = w.masks.as<uint32_t>() + om;
    d.lip = m;
    d.lsp = m + mw[c];
    d.newm = m + 2 * mw[c];
    d.sigarr = m + 3 * mw[c];
    d.signarr = m + 4 * mw[c];
    om += 5 * mw[c];
    d.lis = w.lis.as<node_t>() + ol;
    ol += j.lis_total;
    d.lis_off = j.d_lis_off;
    d.lis_cnt = w.lis_cnt.as<unsigned>() + oc;
    oc += size_t(j.nlis + 1);
    d.nlis = j.nlis;
    d.bits = w.stage.as<uint32_t>() + os;
    d.stage_words = sw[c];
    os += sw[c];
    d.pow2 = j.pow2;
    d.Dx = j.Dx; d.Dy = j.Dy; d.Dz = j.Dz;
    d.nx = j.nx; d.ny = j.ny;
    d.nroots = j.nroots;
    for (int r = 0; r < j.nroots; r++)
      d.roots[r] = j.roots[r];
    d.scr = j.pow2 ? w.scr.as<unsigned long long>() + size_t(c) * kDecThreads * kFastScrPerThread
                   : nullptr;
    max_words = std::max(max_words, sw[c]);
  }
  w.dchunks.reserve(sizeof(DecChunk) * nj);
  rt::h2d(w.dchunks.p, w.h.data(), sizeof(DecChunk) * nj, st);
  const size_t aux_bytes = size_t(nj) * 24;
  w.aux.reserve(aux_bytes);
  unsigned char* aux = w.aux.as<unsigned char>();
  rt::h2d(aux, srcs.data(), nj * 8, st);
  rt::h2d(aux + nj * 8, lens.data(), nj * 8, st);
  rt::h2d(aux + nj * 16, words.data(), nj * 8, st);
  DecChunk* dch = w.dchunks.as<DecChunk>();
  {
    rt::ProfScope ps("dec.stage_bits", st);
    const unsigned gx = unsigned(std::min<size_t>((max_words + 255) / 256, 256));
    LAUNCH(k_stage_bits, dim3(gx, nj), dim3(256), 0, st, dch,
           reinterpret_cast<const unsigned char* const*>(aux),
           reinterpret_cast<const unsigned long long*>(aux + nj * 8),
           reinterpret_cast<const unsigned long long*>(aux + nj * 16));
  }
  {
    rt::ProfScope ps("dec.speck_decode", st);
    if (any_slow)
      LAUNCH(k_speck_decode<T>, dim3(nj), dim3(kDecThreads), 0, st, dch, tree);
    if (any_fast) {
#ifndef SPERR_EMUL
      static bool attr_done = false;
      if (!attr_done) {
        RT_CHECK(cudaFuncSetAttribute(k_speck_decode_fast, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      int(sizeof(FastSmem))));
        attr_done = true;
      }
#endif
      LAUNCH(k_speck_decode_fast, dim3(nj), dim3(kDecThreads), sizeof(FastSmem), st, dch);
    }
  }
  rt::d2h(w.h.data(), w.dchunks.p, sizeof(DecChunk) * nj, st);
  rt::sync(st);
  if (std::getenv("SPERR_B200_DECPROF"))
    for (int c = 0; c < nj && c < 2; c++)
      if (w.h[c].pow2 && !w.h[c].skip)
        std::fprintf(stderr,
                     "decprof job %d n=%llu: lip %.2f  windows %.2f (%llu)  chains %.2f  expand %.2f  walker "
                     "%.2f  refine %.2f  Mcycles\n",
                     c, w.h[c].n, w.h[c].prof[0] * 1e-6, w.h[c].prof[1] * 1e-6, w.h[c].prof[6],
                     w.h[c].prof[2] * 1e-6, w.h[c].prof[3] * 1e-6, w.h[c].prof[4] * 1e-6,
                     w.h[c].prof[5] * 1e-6);
  for (int c = 0; c < nj; c++)
    if (w.h[c].err)
      throw std::runtime_error("SPECK decoder: list capacity exceeded (corrupt stream?)");
}

}  // namespace sperr_b200
