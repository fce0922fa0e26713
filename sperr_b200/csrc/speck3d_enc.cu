// SPECK3D integer encoder as data-parallel "count -> scan -> emit" kernels.
//
// What is reproduced (bit-exactly): SPECK_INT<T>::encode and the 3D sorting / refinement passes,
//   /root/reference/src/SPECK_INT.cpp:110-163,310-357
//   /root/reference/src/SPECK3D_INT.cpp:99-212 (m_sorting_pass, m_code_S), :214-326 (partition)
//   /root/reference/src/SPECK3D_INT_ENC.cpp:161-227
// How (not how the reference does it): the significance of any set at plane n only depends on
// p(S) = msb(max magnitude in S), so after one bottom-up pass that stores p(S) and the size D(S)
// of the depth-first expansion S emits in the plane it turns significant, every emitted bit has a
// position that is computable without walking the lists:
//   * LIP and refinement bits are raster-order prefix sums per plane (k_lipref_*),
//   * the LIS part of a plane is an exclusive scan of (1 + D) over the live sets in list order,
//     followed by a top-down expansion in which each set places its children with D (k_expand),
//   * list order itself is (list index descending, position of the bit that tested the set when it
//     was created), i.e. a radix sort key.
#include "speck3d.h"

namespace sperr_b200 {

// ---------------------------------------------------------------------------------------------
// node helpers
// ---------------------------------------------------------------------------------------------

struct NodeGeom {
  int L, Lc;
  unsigned ix, iy, iz;
  unsigned x0, nxc, y0, nyc, z0, nzc;  // child index ranges in level Lc
  unsigned lenx, leny, lenz;
};

__device__ __forceinline__ void node_geom(const ShapeDev& s, node_t nd, NodeGeom& g)
{
  const ShapeHeader* h = s.h;
  g.L = node_level(nd);
  const LevelDesc& lv = h->lv[g.L];
  g.ix = node_ix(nd); g.iy = node_iy(nd); g.iz = node_iz(nd);
  g.lenx = tab_bnd(s, 0, lv.dx, g.ix + 1) - tab_bnd(s, 0, lv.dx, g.ix);
  g.leny = tab_bnd(s, 1, lv.dy, g.iy + 1) - tab_bnd(s, 1, lv.dy, g.iy);
  g.lenz = tab_bnd(s, 2, lv.dz, g.iz + 1) - tab_bnd(s, 2, lv.dz, g.iz);
  g.Lc = lv.child;
  if (g.Lc < 0) {
    g.x0 = g.ix; g.y0 = g.iy; g.z0 = g.iz;
    g.nxc = g.nyc = g.nzc = 1;
    return;
  }
  const LevelDesc& lc = h->lv[g.Lc];
  if (lc.dx == lv.dx) { g.x0 = g.ix; g.nxc = 1; }
  else { g.x0 = tab_child0(s, 0, lv.dx, g.ix); g.nxc = tab_child0(s, 0, lv.dx, g.ix + 1) - g.x0; }
  if (lc.dy == lv.dy) { g.y0 = g.iy; g.nyc = 1; }
  else { g.y0 = tab_child0(s, 1, lv.dy, g.iy); g.nyc = tab_child0(s, 1, lv.dy, g.iy + 1) - g.y0; }
  if (lc.dz == lv.dz) { g.z0 = g.iz; g.nzc = 1; }
  else { g.z0 = tab_child0(s, 2, lv.dz, g.iz); g.nzc = tab_child0(s, 2, lv.dz, g.iz + 1) - g.z0; }
}

// extent of node (level L, indices) along each axis
__device__ __forceinline__ void node_len(const ShapeDev& s, int L, unsigned ix, unsigned iy,
                                         unsigned iz, unsigned& lx, unsigned& ly, unsigned& lz)
{
  const LevelDesc& lv = s.h->lv[L];
  lx = tab_bnd(s, 0, lv.dx, ix + 1) - tab_bnd(s, 0, lv.dx, ix);
  ly = tab_bnd(s, 1, lv.dy, iy + 1) - tab_bnd(s, 1, lv.dy, iy);
  lz = tab_bnd(s, 2, lv.dz, iz + 1) - tab_bnd(s, 2, lv.dz, iz);
}

__device__ __forceinline__ unsigned long long node_raster(const ShapeDev& s, int L, unsigned ix,
                                                          unsigned iy, unsigned iz)
{
  const ShapeHeader* h = s.h;
  const LevelDesc& lv = h->lv[L];
  const unsigned x = tab_bnd(s, 0, lv.dx, ix), y = tab_bnd(s, 1, lv.dy, iy), z = tab_bnd(s, 2, lv.dz, iz);
  return ((unsigned long long)z * h->ny + y) * h->nx + x;
}

__device__ __forceinline__ size_t node_lin(const LevelDesc& lv, unsigned ix, unsigned iy, unsigned iz)
{
  return ((size_t)iz * lv.cy + iy) * lv.cx + ix;
}

__device__ __forceinline__ int node_p(const ShapeDev& s, const ChunkDev& ch, int L, unsigned ix,
                                      unsigned iy, unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  if (L == s.h->leaf_level)
    return ch.pleaf[node_lin(lv, ix, iy, iz)];
  return ch.pyr_p[lv.p_off + node_lin(lv, ix, iy, iz)];
}

__device__ __forceinline__ unsigned node_d(const ShapeDev& s, const ChunkDev& ch, int L, unsigned ix,
                                           unsigned iy, unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  return ch.pyr_d[lv.p_off + node_lin(lv, ix, iy, iz)];
}

__device__ __forceinline__ unsigned node_lis(const ShapeDev& s, int L, unsigned ix, unsigned iy,
                                             unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  return tab_lev(s, 0, lv.dx, ix) + tab_lev(s, 1, lv.dy, iy) + tab_lev(s, 2, lv.dz, iz);
}

// Is the node below one of the initial sets of its own chain? (only matters for wavelet-packet
// shapes, where several chains overlay the same coefficients)
__device__ __forceinline__ bool node_is_real(const ShapeDev& s, int L, unsigned ix, unsigned iy,
                                             unsigned iz)
{
  const ShapeHeader* h = s.h;
  const LevelDesc& lv = h->lv[L];
  const int x = int(tab_bnd(s, 0, lv.dx, ix)), y = int(tab_bnd(s, 1, lv.dy, iy)),
            z = int(tab_bnd(s, 2, lv.dz, iz));
  for (int g = 0; g < h->ngroups; g++) {
    const GroupDesc& gd = h->grp[g];
    if (gd.chain != lv.chain || lv.j < gd.j_root)
      continue;
    const bool in_outer = x < gd.ox && y < gd.oy && z < gd.oz;
    const bool in_inner = x < gd.ix && y < gd.iy && z < gd.iz;
    if (in_outer && !in_inner)
      return true;
  }
  return false;
}

__device__ __forceinline__ void put_bit(uint32_t* words, unsigned long long pos, unsigned bit)
{
  if (bit)
    atomicOr(&words[pos >> 5], 1u << (pos & 31));
}

// ---------------------------------------------------------------------------------------------
// 1. significance pyramid: p (msb of the max), D (expansion size), cmap (creation plane of pixels)
// ---------------------------------------------------------------------------------------------

__global__ void k_pyr_level(const ChunkDev* chunks, const ShapeDev* shapes, const int* ids, int L,
                            int check_real)
{
  const ChunkDev& ch = chunks[ids[blockIdx.y]];
  if (ch.is_const)
    return;
  const ShapeDev s = shapes[ch.shape];
  const LevelDesc& lv = s.h->lv[L];
  const size_t nodes = (size_t)lv.cx * lv.cy * lv.cz;
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nodes)
    return;
  const unsigned ix = unsigned(idx % lv.cx), iy = unsigned((idx / lv.cx) % lv.cy),
                 iz = unsigned(idx / ((size_t)lv.cx * lv.cy));
  NodeGeom g;
  node_geom(s, make_node(L, ix, iy, iz), g);
  if (g.lenx * g.leny * g.lenz == 1) {  // a pixel that persists at this level
    ch.pyr_p[lv.p_off + idx] = int8_t(node_p(s, ch, g.Lc, g.x0, g.y0, g.z0));
    ch.pyr_d[lv.p_off + idx] = 0;
    return;
  }
  int pc[8];
  unsigned dc[8];
  bool pix[8];
  int nch = 0, pmax = -1;
  for (unsigned cz = 0; cz < g.nzc; cz++)
    for (unsigned cy = 0; cy < g.nyc; cy++)
      for (unsigned cx = 0; cx < g.nxc; cx++) {
        unsigned lx, ly, lz;
        node_len(s, g.Lc, g.x0 + cx, g.y0 + cy, g.z0 + cz, lx, ly, lz);
        pix[nch] = (lx * ly * lz == 1);
        pc[nch] = node_p(s, ch, g.Lc, g.x0 + cx, g.y0 + cy, g.z0 + cz);
        dc[nch] = pix[nch] ? 1u : node_d(s, ch, g.Lc, g.x0 + cx, g.y0 + cy, g.z0 + cz);
        pmax = pc[nch] > pmax ? pc[nch] : pmax;
        nch++;
      }
  unsigned D = 0;
  if (pmax >= 0) {
    int sigc = 0;
    for (int k = 0; k < nch; k++) {
      const bool need = sigc != 0 || k != nch - 1;
      D += need ? 1u : 0u;
      if (pc[k] == pmax) {
        D += dc[k];
        sigc++;
      }
    }
  }
  ch.pyr_p[lv.p_off + idx] = int8_t(pmax);
  ch.pyr_d[lv.p_off + idx] = D;
  // pixels born when this set splits enter the LIP at plane pmax
  if (pmax >= 0 && (!check_real || node_is_real(s, L, ix, iy, iz))) {
    int k = 0;
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++)
          if (pix[k])
            ch.cmap[node_raster(s, g.Lc, g.x0 + cx, g.y0 + cy, g.z0 + cz)] = int8_t(pmax);
  }
}

// ---------------------------------------------------------------------------------------------
// 2. LIP / refinement parts: per-plane raster-order counts, scans and emission
// ---------------------------------------------------------------------------------------------

constexpr int kLrBlock = 1024;

__device__ __forceinline__ unsigned long long load_mag(const ChunkDev& ch, unsigned long long i)
{
  return ch.wide ? reinterpret_cast<const unsigned long long*>(ch.mag)[i]
                 : (unsigned long long)reinterpret_cast<const unsigned*>(ch.mag)[i];
}

// counts[((c*2 + part) * maxp + n) * nblk + blk], part 0 = LIP, 1 = refinement
__global__ void k_lipref_count(const ChunkDev* chunks, unsigned* counts, int maxp, unsigned nblk)
{
  __shared__ unsigned s_lip[64], s_ref[64];
  __shared__ int s_cmax;
  const unsigned c = blockIdx.y, blk = blockIdx.x;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || ch.planes == 0 || (unsigned long long)blk * kLrBlock >= ch.n)
    return;
  const unsigned long long i = (unsigned long long)blk * kLrBlock + threadIdx.x;
  const bool valid = i < ch.n;
  const int p = valid ? int(ch.pleaf[i]) : -1;
  const int cm = valid ? int(ch.cmap[i]) : -1;
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
  for (int n = 0; n < cmax; n++) {
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

// One block per (chunk, part, plane) row: exclusive scan over blocks in place, row total to sizes.
__global__ void k_lipref_scan(const ChunkDev* chunks, unsigned* counts, unsigned long long* sizes,
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
__global__ void k_lipref_emit(const ChunkDev* chunks, const unsigned* counts,
                              const unsigned long long* bases, int maxp, unsigned nblk)
{
  __shared__ unsigned s_w[2][32];
  __shared__ int s_cmax;
  const unsigned c = blockIdx.y, blk = blockIdx.x;
  const ChunkDev& ch = chunks[c];
  if (ch.is_const || ch.planes == 0 || (unsigned long long)blk * kLrBlock >= ch.n)
    return;
  const unsigned long long i = (unsigned long long)blk * kLrBlock + threadIdx.x;
  const bool valid = i < ch.n;
  const int p = valid ? int(ch.pleaf[i]) : -1;
  const int cm = valid ? int(ch.cmap[i]) : -1;
  const unsigned long long mag = (valid && p >= 0) ? load_mag(ch, i) : 0;
  const unsigned sgn = valid ? (ch.signs[i >> 5] >> (i & 31)) & 1u : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1;
  if (threadIdx.x == 0)
    s_cmax = -1;
  __syncthreads();
  const int wmax = __reduce_max_sync(0xffffffffu, cm);
  if (lane == 0 && wmax >= 0)
    atomicMax(&s_cmax, wmax);
  __syncthreads();
  const int cmax = s_cmax;
  const int first = ch.last_plane;  // planes below this were never coded
  for (int n = cmax - 1; n >= first; n--) {
    const bool inlip = cm > n && p <= n;
    const bool newsig = inlip && p == n;
    const bool ref = p > n && !(n == ch.last_plane && ch.stop_after_sort);
    const unsigned b0 = __ballot_sync(0xffffffffu, inlip);
    const unsigned b1 = __ballot_sync(0xffffffffu, newsig);
    const unsigned b2 = __ballot_sync(0xffffffffu, ref);
    if (lane == 0) {
      s_w[0][warp] = __popc(b0) + __popc(b1);
      s_w[1][warp] = __popc(b2);
    }
    __syncthreads();
    if (warp < 2) {
      unsigned v = s_w[warp][lane], inc = v;
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o)
          inc += t;
      }
      s_w[warp][lane] = inc - v;
    }
    __syncthreads();
    if (inlip) {
      const unsigned long long pos = bases[(size_t)(c * 2 + 0) * maxp + n] +
                                     counts[((size_t)(c * 2 + 0) * maxp + n) * nblk + blk] +
                                     s_w[0][warp] + __popc(b0 & lt) + __popc(b1 & lt);
      if (newsig) {
        put_bit(ch.spk, pos, 1);
        put_bit(ch.spk, pos + 1, sgn);
      }
    }
    if (ref) {
      const unsigned long long pos = bases[(size_t)(c * 2 + 1) * maxp + n] +
                                     counts[((size_t)(c * 2 + 1) * maxp + n) * nblk + blk] +
                                     s_w[1][warp] + __popc(b2 & lt);
      put_bit(ch.spk, pos, unsigned(mag >> n) & 1u);
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// 3. LIS part, one plane at a time
// ---------------------------------------------------------------------------------------------

// Single block. Sets planes / budgets and seeds the lists with the initial sets.
__global__ void k_enc_init(EncCtx ctx)
{
  __shared__ unsigned long long s_start[kMaxBatchChunks + 1];
  for (int c = threadIdx.x; c < ctx.nchunks; c += blockDim.x) {
    ChunkDev& ch = ctx.chunks[c];
    int P = 0;
    if (!ch.is_const) {
      const ShapeDev s = ctx.shapes[ch.shape];
      const RootDesc& r0 = s.h->roots[0];
      (void)r0;
      // level 0 of chain 0 is a single node covering the whole chunk
      int top = -1;
      for (int l = 0; l < s.h->nlevels; l++)
        if (s.h->lv[l].chain == 0 && s.h->lv[l].j == 0)
          top = l;
      const int pm = top >= 0 ? int(ch.pyr_p[s.h->lv[top].p_off]) : int(ch.pleaf[0]);
      P = pm + 1;
    }
    ch.planes = P;
    ch.active = P > 0;
    ch.cursor = 0;
    ch.total_bits = 0;
    ch.last_plane = P > 0 ? 0 : 0;
    ch.stop_after_sort = 0;
    ch.ncand = 0;
    s_start[c] = P > 0 ? ctx.shapes[ch.shape].h->nroots : 0;
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
    const ShapeDev s = ctx.shapes[ch.shape];
    for (int r = 0; r < s.h->nroots; r++) {
      const RootDesc& rd = s.h->roots[r];
      ctx.rkey[s_start[c] + r] = make_key(c, unsigned(s.h->nlis - 1 - rd.lis), unsigned(rd.order));
      ctx.rnode[s_start[c] + r] = make_node(rd.level, rd.ix, rd.iy, rd.iz);
    }
  }
}

// step s of the bit-plane loop: which plane does each chunk code?
__global__ void k_plane_pre(EncCtx ctx, int step)
{
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= ctx.nchunks)
    return;
  ChunkDev& ch = ctx.chunks[c];
  ch.cur_n = ch.planes - 1 - step;
  ch.plane_live = ch.active && ch.cur_n >= 0;
  ch.ncand = 0;
}

__global__ void k_root_seglen(EncCtx ctx, unsigned long long nroots)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nroots;
       i += stride) {
    const ChunkDev& ch = ctx.chunks[key_chunk(ctx.rkey[i])];
    const ShapeDev s = ctx.shapes[ch.shape];
    const node_t nd = ctx.rnode[i];
    const int L = node_level(nd);
    const int p = node_p(s, ch, L, node_ix(nd), node_iy(nd), node_iz(nd));
    unsigned seg = 1;
    if (p == ch.cur_n)
      seg += node_d(s, ch, L, node_ix(nd), node_iy(nd), node_iz(nd));
    ctx.rseg[i] = seg;
  }
}

// Places the three parts of this plane and applies the fixed-rate budget rules
// (src/SPECK_INT.cpp:145-158).
__global__ void k_plane_begin(EncCtx ctx)
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

struct Appender {
  EncCtx ctx;
  __device__ __forceinline__ void cand(unsigned c, unsigned long long key, node_t nd) const
  {
    const unsigned long long slot = atomicAdd(ctx.cand_count, 1ull);
    if (slot >= ctx.cand_cap) {
      atomicOr(ctx.err, 1u);
      return;
    }
    ctx.ckey[slot] = key;
    ctx.cnode[slot] = nd;
    atomicAdd(&ctx.chunks[c].ncand, 1u);
  }
  __device__ __forceinline__ void front(int dst, unsigned c, node_t nd, unsigned long long pos) const
  {
    const unsigned long long slot = atomicAdd(&ctx.fcount[dst], 1ull);
    if (slot >= ctx.front_cap) {
      atomicOr(ctx.err, 2u);
      return;
    }
    ctx.fnode[dst][slot] = nd;
    ctx.fpos[dst][slot] = (pos << 10) | c;
  }
};

__global__ void k_root_emit(EncCtx ctx, unsigned long long nroots)
{
  const Appender app{ctx};
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < nroots;
       i += stride) {
    const unsigned long long key = ctx.rkey[i];
    const unsigned c = key_chunk(key);
    const ChunkDev& ch = ctx.chunks[c];
    if (!ch.plane_live)
      continue;
    const unsigned long long pos = ch.lis_base + (ctx.rpos[i] - ctx.rpos[ctx.rstart[c]]);
    const node_t nd = ctx.rnode[i];
    const ShapeDev s = ctx.shapes[ch.shape];
    if (node_p(s, ch, node_level(nd), node_ix(nd), node_iy(nd), node_iz(nd)) == ch.cur_n) {
      put_bit(ch.spk, pos, 1);
      app.front(0, c, nd, pos + 1);
    }
    else if (ch.next_needed)  // insignificant: stays in its list, same position
      app.cand(c, key, nd);
  }
}

// Emits the tokens of the pixel children of a set whose children are all pixels.
__device__ __forceinline__ unsigned long long expand_leafparent(const ShapeDev& s, const ChunkDev& ch,
                                                                int Lc, unsigned x0, unsigned nxc,
                                                                unsigned y0, unsigned nyc, unsigned z0,
                                                                unsigned nzc, int n,
                                                                unsigned long long cur)
{
  const int nch = int(nxc * nyc * nzc);
  int k = 0, sigc = 0;
  for (unsigned cz = 0; cz < nzc; cz++)
    for (unsigned cy = 0; cy < nyc; cy++)
      for (unsigned cx = 0; cx < nxc; cx++, k++) {
        const bool need = sigc != 0 || k != nch - 1;
        const bool sig = !need || node_p(s, ch, Lc, x0 + cx, y0 + cy, z0 + cz) == n;
        if (need) {
          put_bit(ch.spk, cur, sig);
          cur++;
        }
        if (sig) {
          const unsigned long long r = node_raster(s, Lc, x0 + cx, y0 + cy, z0 + cz);
          put_bit(ch.spk, cur, (ch.signs[r >> 5] >> (r & 31)) & 1u);
          cur++;
          sigc++;
        }
      }
  return cur;
}

// One thread per significant set: m_code_S (src/SPECK3D_INT.cpp:140-212) with every child placed
// by its D instead of by recursion order.
__global__ void k_expand(EncCtx ctx, int src)
{
  const Appender app{ctx};
  const int dst = src ^ 1;
  const unsigned long long count = ctx.fcount[src];
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count;
       i += stride) {
    const node_t nd = ctx.fnode[src][i];
    const unsigned long long pc = ctx.fpos[src][i];
    const unsigned c = unsigned(pc & 1023u);
    unsigned long long cur = pc >> 10;
    const ChunkDev& ch = ctx.chunks[c];
    const ShapeDev s = ctx.shapes[ch.shape];
    const int n = ch.cur_n;
    NodeGeom g;
    node_geom(s, nd, g);
    const int nch = int(g.nxc * g.nyc * g.nzc);
    int k = 0, sigc = 0;
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
          const unsigned jx = g.x0 + cx, jy = g.y0 + cy, jz = g.z0 + cz;
          const bool need = sigc != 0 || k != nch - 1;
          const bool sig = !need || node_p(s, ch, g.Lc, jx, jy, jz) == n;
          if (need) {
            put_bit(ch.spk, cur, sig);
            cur++;
          }
          unsigned lx, ly, lz;
          node_len(s, g.Lc, jx, jy, jz, lx, ly, lz);
          if (lx * ly * lz == 1) {
            if (sig) {
              const unsigned long long r = node_raster(s, g.Lc, jx, jy, jz);
              put_bit(ch.spk, cur, (ch.signs[r >> 5] >> (r & 31)) & 1u);
              cur++;
              sigc++;
            }
          }
          else if (sig) {
            sigc++;
            if (lx <= 2 && ly <= 2 && lz <= 2) {
              NodeGeom cg;
              node_geom(s, make_node(g.Lc, jx, jy, jz), cg);
              cur = expand_leafparent(s, ch, cg.Lc, cg.x0, cg.nxc, cg.y0, cg.nyc, cg.z0, cg.nzc, n, cur);
            }
            else {
              app.front(dst, c, make_node(g.Lc, jx, jy, jz), cur);
              cur += node_d(s, ch, g.Lc, jx, jy, jz);
            }
          }
          else if (ch.next_needed) {
            const unsigned lis = node_lis(s, g.Lc, jx, jy, jz);
            app.cand(c, make_key(c, unsigned(s.h->nlis - 1) - lis, (cur - 1) + 64),
                     make_node(g.Lc, jx, jy, jz));
          }
        }
  }
}

__global__ void k_front_reset(EncCtx ctx, int which)
{
  if (threadIdx.x == 0 && blockIdx.x == 0)
    ctx.fcount[which] = 0;
}

// Single block: list segment of every chunk for the next plane.
__global__ void k_plane_end(EncCtx ctx)
{
  __shared__ unsigned long long s_start[kMaxBatchChunks + 1];
  for (int c = threadIdx.x; c < ctx.nchunks; c += blockDim.x)
    s_start[c] = ctx.chunks[c].ncand;
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
    *ctx.cand_count = 0;
    ctx.fcount[0] = 0;
    ctx.fcount[1] = 0;
  }
  __syncthreads();
  for (int c = threadIdx.x; c <= ctx.nchunks; c += blockDim.x)
    ctx.rstart[c] = s_start[c];
}

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------

void Speck3DEncoder::encode(ChunkDev* d_chunks, const std::vector<ChunkDev>& h_chunks_in,
                            const ShapeDev* d_shapes, const std::vector<ShapeTables>& shapes,
                            std::vector<EncResult>& results, cudaStream_t st)
{
  const int nchunks = int(h_chunks_in.size());
  if (nchunks > kMaxBatchChunks)
    throw std::runtime_error("too many chunks in one batch");
  results.assign(nchunks, EncResult());

  // ---- 1. pyramid, per shape group (levels of one chain are built child -> parent) ----
  size_t max_n = 0;
  unsigned long long cap_nodes = 0;
  for (auto& c : h_chunks_in) {
    max_n = std::max<size_t>(max_n, c.n);
    cap_nodes += shapes[c.shape].h.set_nodes;
  }
  std::vector<std::vector<int>> groups(shapes.size());
  for (int c = 0; c < nchunks; c++)
    groups[h_chunks_in[c].shape].push_back(c);
  ids_.reserve(nchunks * sizeof(int));
  {
    std::vector<int> flat;
    for (auto& g : groups)
      flat.insert(flat.end(), g.begin(), g.end());
    rt::h2d(ids_.p, flat.data(), flat.size() * sizeof(int), st);
  }
  size_t id_off = 0;
  for (size_t si = 0; si < shapes.size(); si++) {
    const auto& g = groups[si];
    if (g.empty())
      continue;
    const ShapeHeader& h = shapes[si].h;
    const int* d_ids = ids_.as<int>() + id_off;
    id_off += g.size();
    // order: deepest levels first. Within a chain j descending works because child = j + 1.
    std::vector<int> order;
    for (int l = 0; l < h.nlevels; l++)
      if (l != h.leaf_level)
        order.push_back(l);
    std::sort(order.begin(), order.end(), [&](int a, int b) { return h.lv[a].j > h.lv[b].j; });
    const int check_real = h.dyadic < 0 ? 1 : 0;
    for (int l : order) {
      const size_t nodes = (size_t)h.lv[l].cx * h.lv[l].cy * h.lv[l].cz;
      LAUNCH(k_pyr_level, dim3(unsigned((nodes + 255) / 256), unsigned(g.size())), dim3(256), 0, st,
             d_chunks, d_shapes, d_ids, l, check_real);
    }
  }

  // ---- 2. context ----
  EncCtx ctx;
  ctx.chunks = d_chunks;
  ctx.nchunks = nchunks;
  ctx.shapes = d_shapes;
  ctx.cand_cap = cap_nodes + 64ull * nchunks;
  ctx.front_cap = cap_nodes + 64ull * nchunks;
  const size_t cap = size_t(ctx.cand_cap);
  keys_[0].reserve(cap * 8); keys_[1].reserve(cap * 8);
  nodes_[0].reserve(cap * 8); nodes_[1].reserve(cap * 8);
  fnode_[0].reserve(cap * 8); fnode_[1].reserve(cap * 8);
  fpos_[0].reserve(cap * 8); fpos_[1].reserve(cap * 8);
  rseg_.reserve((cap + 1) * 4);
  rpos_.reserve((cap + 2) * 8);
  scan_tmp_.reserve(scan_tmp_bytes(cap + 1));
  const size_t sort_bytes = sort_tmp_bytes(cap);
  sort_tmp_.reserve(sort_bytes);
  small_.reserve(4096 + (nchunks + 1) * 8);
  rt::dset(small_.p, 0, small_.bytes, st);
  unsigned long long* small = small_.as<unsigned long long>();
  ctx.total_roots = small + 0;
  ctx.cand_count = small + 1;
  ctx.fcount = small + 2;  // two counters
  ctx.err = reinterpret_cast<unsigned*>(small + 4);
  ctx.rstart = small + 8;
  ctx.rkey = keys_[0].as<unsigned long long>();
  ctx.rnode = nodes_[0].as<node_t>();
  ctx.ckey = keys_[1].as<unsigned long long>();
  ctx.cnode = nodes_[1].as<node_t>();
  ctx.fnode[0] = fnode_[0].as<node_t>(); ctx.fnode[1] = fnode_[1].as<node_t>();
  ctx.fpos[0] = fpos_[0].as<unsigned long long>(); ctx.fpos[1] = fpos_[1].as<unsigned long long>();
  ctx.rseg = rseg_.as<unsigned>();
  ctx.rpos = rpos_.as<unsigned long long>();

  LAUNCH(k_enc_init, dim3(1), dim3(1024), 0, st, ctx);

  // planes are only known on the device: fetch them to size the staging buffers and tables
  std::vector<ChunkDev> hc(nchunks);
  rt::d2h(hc.data(), d_chunks, sizeof(ChunkDev) * nchunks, st);
  unsigned long long total_roots = 0;
  rt::d2h(&total_roots, ctx.total_roots, 8, st);
  rt::sync(st);
  int maxp = 1;
  for (auto& c : hc)
    maxp = std::max(maxp, c.planes);
  ctx.maxp = maxp;

  // staging bit arrays
  {
    std::vector<unsigned long long> words_off(nchunks + 1, 0);
    for (int c = 0; c < nchunks; c++) {
      const ShapeHeader& h = shapes[hc[c].shape].h;
      unsigned long long bits = 0;
      if (hc[c].planes > 0) {
        bits = hc[c].n * (unsigned long long)(hc[c].planes + 1) +
               h.set_nodes * (unsigned long long)hc[c].planes + 64;
        if (hc[c].budget != ~0ull)
          bits = std::min(bits, hc[c].budget + 3 * hc[c].n + h.set_nodes + 64);
      }
      words_off[c + 1] = words_off[c] + (bits + 31) / 32 + 2;
    }
    stage_.reserve(words_off[nchunks] * 4);
    rt::dset(stage_.p, 0, words_off[nchunks] * 4, st);
    for (int c = 0; c < nchunks; c++) {
      hc[c].spk = stage_.as<uint32_t>() + words_off[c];
      hc[c].spk_cap_bits = (words_off[c + 1] - words_off[c]) * 32;
    }
    rt::h2d(d_chunks, hc.data(), sizeof(ChunkDev) * nchunks, st);
  }

  // ---- 3. LIP / refinement counts for every plane ----
  const unsigned nblk = unsigned((max_n + kLrBlock - 1) / kLrBlock);
  const size_t ncounts = (size_t)nchunks * 2 * maxp * nblk;
  counts_.reserve(ncounts * 4);
  rt::dset(counts_.p, 0, ncounts * 4, st);
  sizes_.reserve((size_t)nchunks * 2 * maxp * 8 * 2);
  rt::dset(sizes_.p, 0, (size_t)nchunks * 2 * maxp * 8 * 2, st);
  ctx.sizes = sizes_.as<unsigned long long>();
  ctx.bases = ctx.sizes + (size_t)nchunks * 2 * maxp;
  LAUNCH(k_lipref_count, dim3(nblk, nchunks), dim3(kLrBlock), 0, st, d_chunks, counts_.as<unsigned>(),
         maxp, nblk);
  LAUNCH(k_lipref_scan, dim3(2 * maxp, nchunks), dim3(1024), 0, st, d_chunks, counts_.as<unsigned>(),
         ctx.sizes, maxp, nblk);

  // ---- 4. bit-plane loop (LIS part) ----
  int max_chain = 1;
  for (auto& sh : shapes)
    for (int l = 0; l < sh.h.nlevels; l++)
      max_chain = std::max(max_chain, sh.h.lv[l].j + 2);
  const unsigned cgrid = unsigned((nchunks + 127) / 128);
  for (int step = 0; step < maxp; step++) {
    LAUNCH(k_plane_pre, dim3(cgrid), dim3(128), 0, st, ctx, step);
    const unsigned rgrid = unsigned(std::min<unsigned long long>((total_roots + 255) / 256, 148 * 16));
    if (total_roots) {
      LAUNCH(k_root_seglen, dim3(rgrid), dim3(256), 0, st, ctx, total_roots);
    }
    exclusive_scan_u32(ctx.rseg, ctx.rpos, total_roots, scan_tmp_.p, st);
    LAUNCH(k_plane_begin, dim3(cgrid), dim3(128), 0, st, ctx);
    if (total_roots) {
      LAUNCH(k_root_emit, dim3(rgrid), dim3(256), 0, st, ctx, total_roots);
      for (int d = 0; d < max_chain; d++) {
        LAUNCH(k_expand, dim3(148 * 8), dim3(256), 0, st, ctx, d & 1);
        LAUNCH(k_front_reset, dim3(1), dim3(32), 0, st, ctx, d & 1);
      }
    }
    LAUNCH(k_plane_end, dim3(1), dim3(1024), 0, st, ctx);
    rt::d2h(&total_roots, ctx.total_roots, 8, st);
    rt::sync(st);
    if (total_roots == 0 && step + 1 < maxp) {
      // nothing left in any list: the remaining planes have no LIS part, but k_plane_begin must
      // still place their LIP / refinement parts
      continue;
    }
    if (total_roots > ctx.cand_cap)
      throw std::runtime_error("SPECK list overflow");
    sort_pairs_u64(ctx.ckey, ctx.rkey, ctx.cnode, ctx.rnode, total_roots, kKeyBits, sort_tmp_.p,
                   sort_bytes, st);
  }

  // ---- 5. LIP / refinement emission ----
  LAUNCH(k_lipref_emit, dim3(nblk, nchunks), dim3(kLrBlock), 0, st, d_chunks, counts_.as<unsigned>(),
         ctx.bases, maxp, nblk);

  rt::d2h(hc.data(), d_chunks, sizeof(ChunkDev) * nchunks, st);
  unsigned err = 0;
  rt::d2h(&err, ctx.err, 4, st);
  rt::sync(st);
  if (err)
    throw std::runtime_error("SPECK encoder work-list overflow");
  for (int c = 0; c < nchunks; c++) {
    results[c].planes = hc[c].planes;
    results[c].total_bits = hc[c].total_bits;
    results[c].payload = hc[c].spk;
    const unsigned long long pack = std::min(hc[c].total_bits, hc[c].budget);
    results[c].payload_bytes = size_t((pack + 7) / 8);
  }
}

}  // namespace sperr_b200
