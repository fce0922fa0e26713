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
#include "speck_engine.cuh"
#include "tree3d.cuh"

namespace sperr_b200 {

__device__ __forceinline__ int node_p(const ShapeDev& s, const ChunkDev& ch, int L, unsigned ix,
                                      unsigned iy, unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  if (L == s.h->leaf_level)
    return gptr(ch.pleaf)[node_lin(lv, ix, iy, iz)];
  return gptr(ch.pyr_p)[lv.p_off + node_lin(lv, ix, iy, iz)];
}

__device__ __forceinline__ unsigned node_d(const ShapeDev& s, const ChunkDev& ch, int L, unsigned ix,
                                           unsigned iy, unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  return gptr(ch.pyr_d)[lv.p_off + node_lin(lv, ix, iy, iz)];
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

// ---------------------------------------------------------------------------------------------
// 1. significance pyramid: p (msb of the max), D (expansion size), cmap (creation plane of pixels)
// ---------------------------------------------------------------------------------------------

// rev: the children are coded in reverse raster order (2D slices: BR, BL, TR, TL,
// /root/reference/src/SPECK2D_INT.cpp:109-148)
__global__ void k_pyr_level(const ChunkDev* chunks, const ShapeDev* shapes, const int* ids, int L,
                            int check_real, int rev)
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
    gptr(ch.pyr_p)[lv.p_off + idx] = int8_t(node_p(s, ch, g.Lc, g.x0, g.y0, g.z0));
    gptr(ch.pyr_d)[lv.p_off + idx] = 0;
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
      const int kk = rev ? nch - 1 - k : k;
      const bool need = sigc != 0 || k != nch - 1;
      D += need ? 1u : 0u;
      if (pc[kk] == pmax) {
        D += dc[kk];
        sigc++;
      }
    }
  }
  gptr(ch.pyr_p)[lv.p_off + idx] = int8_t(pmax);
  gptr(ch.pyr_d)[lv.p_off + idx] = D;
  // pixels born when this set splits enter the LIP at plane pmax
  if (pmax >= 0 && (!check_real || node_is_real(s, L, ix, iy, iz))) {
    int k = 0;
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++)
          if (pix[k])
            gptr(ch.cmap)[node_raster(s, g.Lc, g.x0 + cx, g.y0 + cy, g.z0 + cz)] = int8_t(pmax);
  }
}


// ---------------------------------------------------------------------------------------------
// 2. tree policy of the 3D coder: implicit octree addressed through the shape tables
// ---------------------------------------------------------------------------------------------

struct Tree3D {
  static constexpr bool kHasLeaf8 = false;
  static constexpr bool kIsOutlierTree = false;
  struct Data {
    const ShapeDev* shapes;
  };

  static __device__ __forceinline__ void pd(const Data& t, const ChunkDev& ch, unsigned, node_t nd,
                                            int& p, unsigned& d)
  {
    const ShapeDev s = t.shapes[ch.shape];
    const int L = node_level(nd);
    p = node_p(s, ch, L, node_ix(nd), node_iy(nd), node_iz(nd));
    d = L == s.h->leaf_level ? 0u : node_d(s, ch, L, node_ix(nd), node_iy(nd), node_iz(nd));
  }

  static __device__ __forceinline__ int children(const Data& t, const ChunkDev& ch, unsigned,
                                                 node_t nd, ChildRec* out)
  {
    const ShapeDev s = t.shapes[ch.shape];
    NodeGeom g;
    node_geom(s, nd, g);
    int k = 0;
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
          const unsigned jx = g.x0 + cx, jy = g.y0 + cy, jz = g.z0 + cz;
          ChildRec& r = out[k];
          r.id = make_node(g.Lc, jx, jy, jz);
          r.p = node_p(s, ch, g.Lc, jx, jy, jz);
          unsigned lx, ly, lz;
          node_len(s, g.Lc, jx, jy, jz, lx, ly, lz);
          if (lx * ly * lz == 1) {
            r.kind = 0;
            r.d = 0;
            r.lis_desc = 0;
            const unsigned long long ri = node_raster(s, g.Lc, jx, jy, jz);
            r.sign = (gptr(ch.signs)[ri >> 5] >> (ri & 31)) & 1u;
          }
          else {
            r.kind = (lx <= 2 && ly <= 2 && lz <= 2) ? 1 : 2;
            r.d = node_d(s, ch, g.Lc, jx, jy, jz);
            r.lis_desc = unsigned(s.h->nlis - 1) - node_lis(s, g.Lc, jx, jy, jz);
            r.sign = 0;
          }
        }
    return k;
  }

  static __device__ __forceinline__ int planes(const Data& t, const ChunkDev& ch, unsigned)
  {
    const ShapeDev s = t.shapes[ch.shape];
    // level 0 of chain 0 is a single node covering the whole chunk
    int top = -1;
    for (int l = 0; l < s.h->nlevels; l++)
      if (s.h->lv[l].chain == 0 && s.h->lv[l].j == 0)
        top = l;
    const int pm = top >= 0 ? int(gptr(ch.pyr_p)[s.h->lv[top].p_off]) : int(gptr(ch.pleaf)[0]);
    return pm + 1;
  }

  static __device__ __forceinline__ int num_roots(const Data& t, const ChunkDev& ch, unsigned)
  {
    return t.shapes[ch.shape].h->nroots;
  }

  static __device__ __forceinline__ void root(const Data& t, const ChunkDev& ch, unsigned, int r,
                                              node_t& nd, unsigned& lis_desc, unsigned& order)
  {
    const ShapeHeader* h = t.shapes[ch.shape].h;
    const RootDesc& rd = h->roots[r];
    nd = make_node(rd.level, rd.ix, rd.iy, rd.iz);
    lis_desc = unsigned(h->nlis - 1 - rd.lis);
    order = unsigned(rd.order);
  }
};

// ---------------------------------------------------------------------------------------------
// 2b. the same tree for power-of-two dyadic chunks (the bench shape: 256^3): every set at depth j
//     is an aligned box, so children, pyramid slots and list indices are shifts instead of table
//     look-ups. Node ids carry the depth j (not the LevelDesc index): make_node(j, ix, iy, iz).
// ---------------------------------------------------------------------------------------------

struct Pow2Info {
  int Dx, Dy, Dz, J, nlis;
  unsigned nx, ny;
  unsigned long long p_off[kMaxAxisDepth + 1];   // pyramid offset of depth j < J
};

__device__ __forceinline__ size_t pow2_lin(const Pow2Info& g, int j, unsigned ix, unsigned iy, unsigned iz)
{
  const int bx = min(j, g.Dx), by = min(j, g.Dy);
  return ((size_t)iz << (bx + by)) | ((size_t)iy << bx) | ix;
}

__global__ void k_pyr_pow2(const ChunkDev* chunks, Pow2Info g, int j)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  if (ch.is_const)
    return;
  const int bx = min(j, g.Dx), by = min(j, g.Dy), bz = min(j, g.Dz);
  const size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >> (bx + by + bz))
    return;
  const unsigned ix = unsigned(idx) & ((1u << bx) - 1u), iy = unsigned(idx >> bx) & ((1u << by) - 1u),
                 iz = unsigned(idx >> (bx + by));
  const int sx = j < g.Dx, sy = j < g.Dy, sz = j < g.Dz;
  const int nch = 1 << (sx + sy + sz), cj = j + 1;
  const bool leaf = cj == g.J;
  int pc[8];
  unsigned dc[8];
  size_t cpos[8];
  int pmax = -1;
  for (int k = 0; k < nch; k++) {
    const unsigned jx = sx ? ix * 2 + (unsigned(k) & 1u) : ix;
    const unsigned jy = sy ? iy * 2 + ((unsigned(k) >> sx) & 1u) : iy;
    const unsigned jz = sz ? iz * 2 + ((unsigned(k) >> (sx + sy)) & 1u) : iz;
    if (leaf) {
      cpos[k] = ((size_t)jz * g.ny + jy) * g.nx + jx;
      pc[k] = gptr(ch.pleaf)[cpos[k]];
      dc[k] = 1u;
    }
    else {
      cpos[k] = g.p_off[cj] + pow2_lin(g, cj, jx, jy, jz);
      pc[k] = gptr(ch.pyr_p)[cpos[k]];
      dc[k] = gptr(ch.pyr_d)[cpos[k]];
    }
    pmax = pc[k] > pmax ? pc[k] : pmax;
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
  gptr(ch.pyr_p)[g.p_off[j] + idx] = int8_t(pmax);
  gptr(ch.pyr_d)[g.p_off[j] + idx] = D;
  if (leaf && pmax >= 0)   // pixels born when this set splits enter the LIP at plane pmax
    for (int k = 0; k < nch; k++)
      gptr(ch.cmap)[cpos[k]] = int8_t(pmax);
}

struct Tree3DPow2 {
  static constexpr bool kHasLeaf8 = true;
  static constexpr bool kIsOutlierTree = false;
  struct Data {
    const ShapeDev* shapes;
    Pow2Info g;
  };

  // the single coefficients below a set of depth J - 1, in coding order (x fastest): pairs along x
  // share a 16-bit load of the msb map and a word of the sign map
  static __device__ __forceinline__ int leaf8(const Data& t, const ChunkDev& ch, node_t nd, int* p,
                                              unsigned& sg)
  {
    const Pow2Info& g = t.g;
    const int j = node_level(nd);
    const unsigned ix = node_ix(nd), iy = node_iy(nd), iz = node_iz(nd);
    const int sx = j < g.Dx, sy = j < g.Dy, sz = j < g.Dz;
    const int rows = 1 << (sy + sz);
    sg = 0;
    int k = 0;
    for (int r = 0; r < rows; r++) {
      const unsigned jy = sy ? iy * 2 + (unsigned(r) & 1u) : iy;
      const unsigned jz = sz ? iz * 2 + ((unsigned(r) >> sy) & 1u) : iz;
      const size_t ri = ((size_t)jz * g.ny + jy) * g.nx + (sx ? ix * 2 : ix);
      if (sx) {   // ri is even: both values in one aligned 16-bit word, both sign bits in one word
        const unsigned v = *reinterpret_cast<const unsigned short*>(gptr(ch.pleaf) + ri);
        p[k] = int(int8_t(v & 0xff));
        p[k + 1] = int(int8_t(v >> 8));
        sg |= ((gptr(ch.signs)[ri >> 5] >> (ri & 31)) & 3u) << k;
        k += 2;
      }
      else {
        p[k] = gptr(ch.pleaf)[ri];
        sg |= ((gptr(ch.signs)[ri >> 5] >> (ri & 31)) & 1u) << k;
        k += 1;
      }
    }
    return k;
  }

  static __device__ __forceinline__ void pd(const Data& t, const ChunkDev& ch, unsigned, node_t nd,
                                            int& p, unsigned& d)
  {
    const Pow2Info& g = t.g;
    const int j = node_level(nd);
    if (j == g.J) {
      p = gptr(ch.pleaf)[((size_t)node_iz(nd) * g.ny + node_iy(nd)) * g.nx + node_ix(nd)];
      d = 0;
      return;
    }
    const size_t at = g.p_off[j] + pow2_lin(g, j, node_ix(nd), node_iy(nd), node_iz(nd));
    p = gptr(ch.pyr_p)[at];
    d = gptr(ch.pyr_d)[at];
  }

  static __device__ __forceinline__ int children(const Data& t, const ChunkDev& ch, unsigned,
                                                 node_t nd, ChildRec* out)
  {
    const Pow2Info& g = t.g;
    const int j = node_level(nd), cj = j + 1;
    const unsigned ix = node_ix(nd), iy = node_iy(nd), iz = node_iz(nd);
    const int sx = j < g.Dx, sy = j < g.Dy, sz = j < g.Dz;
    const int nch = 1 << (sx + sy + sz);
    const int kind = cj == g.J ? 0 : (cj == g.J - 1 ? 1 : 2);
    const unsigned lis_desc = unsigned(g.nlis - 1 - (min(cj, g.Dx) + min(cj, g.Dy) + min(cj, g.Dz)));
    for (int k = 0; k < nch; k++) {
      const unsigned jx = sx ? ix * 2 + (unsigned(k) & 1u) : ix;
      const unsigned jy = sy ? iy * 2 + ((unsigned(k) >> sx) & 1u) : iy;
      const unsigned jz = sz ? iz * 2 + ((unsigned(k) >> (sx + sy)) & 1u) : iz;
      ChildRec& r = out[k];
      r.id = make_node(cj, jx, jy, jz);
      r.kind = kind;
      if (kind == 0) {
        const size_t ri = ((size_t)jz * g.ny + jy) * g.nx + jx;
        r.p = gptr(ch.pleaf)[ri];
        r.d = 0;
        r.lis_desc = 0;
        r.sign = (gptr(ch.signs)[ri >> 5] >> (ri & 31)) & 1u;
      }
      else {
        const size_t at = g.p_off[cj] + pow2_lin(g, cj, jx, jy, jz);
        r.p = gptr(ch.pyr_p)[at];
        r.d = gptr(ch.pyr_d)[at];
        r.lis_desc = lis_desc;
        r.sign = 0;
      }
    }
    return nch;
  }

  static __device__ __forceinline__ int planes(const Data& t, const ChunkDev& ch, unsigned)
  {
    return int(gptr(ch.pyr_p)[t.g.p_off[0]]) + 1;
  }

  static __device__ __forceinline__ int num_roots(const Data& t, const ChunkDev& ch, unsigned)
  {
    return t.shapes[ch.shape].h->nroots;
  }

  static __device__ __forceinline__ void root(const Data& t, const ChunkDev& ch, unsigned, int r,
                                              node_t& nd, unsigned& lis_desc, unsigned& order)
  {
    const ShapeHeader* h = t.shapes[ch.shape].h;
    const RootDesc& rd = h->roots[r];
    nd = make_node(h->lv[rd.level].j, rd.ix, rd.iy, rd.iz);
    lis_desc = unsigned(h->nlis - 1 - rd.lis);
    order = unsigned(rd.order);
  }
};

// ---------------------------------------------------------------------------------------------
// 3. tree policy of the 2D coder (SPECK2D_INT, /root/reference/src/SPECK2D_INT.cpp:10-218,
//    src/SPECK2D_INT_ENC.cpp:7-121): quadtree S sets coded in reverse raster order, plus the set
//    I_l = everything outside the approximation band of transform level l. I_l splits into
//    BR, TR, BL of level l (always tested) and I_(l-1) (implied significant when none of the three
//    was), and it is tested after all lists: it sorts behind every S set.
//    I_l's p / D live behind the pyramid: pyr_p[pyr_nodes + l], pyr_d[pyr_nodes + l].
// ---------------------------------------------------------------------------------------------

constexpr int kINodeLevel = 0xFF;
__device__ __forceinline__ bool is_inode(node_t nd) { return node_level(nd) == kINodeLevel; }

// one thread per chunk: I_1 .. I_nxf bottom-up
__global__ void k_pyr_iset(const ChunkDev* chunks, const ShapeDev* shapes, const int* ids, int nids)
{
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nids)
    return;
  const ChunkDev& ch = chunks[ids[t]];
  if (ch.is_const)
    return;
  const ShapeDev s = shapes[ch.shape];
  const ShapeHeader* h = s.h;
  int pi = -1;        // p of I_(l-1); I_0 is empty
  unsigned di = 0;
  for (int l = 1; l <= h->nxf2d; l++) {
    const int L = h->lv2d[l];
    const int cx[3] = {1, 1, 0}, cy[3] = {1, 0, 1};   // BR, TR, BL (src/SPECK2D_INT.cpp:150-185)
    int pc[3], pmax = pi;
    unsigned dc[3];
    for (int k = 0; k < 3; k++) {
      pc[k] = node_p(s, ch, L, cx[k], cy[k], 0);
      dc[k] = node_d(s, ch, L, cx[k], cy[k], 0);
      pmax = pc[k] > pmax ? pc[k] : pmax;
    }
    unsigned D = 0;
    if (pmax >= 0) {
      int sigc = 0;
      for (int k = 0; k < 3; k++) {
        D += 1;
        if (pc[k] == pmax) {
          D += dc[k];
          sigc++;
        }
      }
      if (l > 1) {
        D += sigc != 0 ? 1u : 0u;
        if (pi == pmax)
          D += di;
      }
    }
    gptr(ch.pyr_p)[h->pyr_nodes + l] = int8_t(pmax);
    gptr(ch.pyr_d)[h->pyr_nodes + l] = D;
    pi = pmax;
    di = D;
  }
}

struct Tree2D {
  static constexpr bool kHasLeaf8 = false;
  static constexpr bool kIsOutlierTree = false;
  typedef Tree3D::Data Data;

  static __device__ __forceinline__ void pd(const Data& t, const ChunkDev& ch, unsigned c, node_t nd,
                                            int& p, unsigned& d)
  {
    if (is_inode(nd)) {
      const ShapeHeader* h = t.shapes[ch.shape].h;
      p = gptr(ch.pyr_p)[h->pyr_nodes + node_ix(nd)];
      d = gptr(ch.pyr_d)[h->pyr_nodes + node_ix(nd)];
      return;
    }
    Tree3D::pd(t, ch, c, nd, p, d);
  }

  static __device__ __forceinline__ void fill_set(const ShapeDev& s, const ChunkDev& ch, int L,
                                                  unsigned jx, unsigned jy, ChildRec& r)
  {
    r.id = make_node(L, jx, jy, 0);
    r.p = node_p(s, ch, L, jx, jy, 0);
    unsigned lx, ly, lz;
    node_len(s, L, jx, jy, 0, lx, ly, lz);
    r.kind = (lx <= 2 && ly <= 2) ? 1 : 2;
    r.d = node_d(s, ch, L, jx, jy, 0);
    r.lis_desc = unsigned(s.h->nlis - 1 - s.h->lv[L].j);   // part_level = chain position
    r.sign = 0;
  }

  // bit 8 of the result: every child is tested explicitly (no implied last child)
  static __device__ __forceinline__ int children(const Data& t, const ChunkDev& ch, unsigned c,
                                                 node_t nd, ChildRec* out)
  {
    const ShapeDev s = t.shapes[ch.shape];
    const ShapeHeader* h = s.h;
    if (is_inode(nd)) {
      const int l = int(node_ix(nd));
      const int L = h->lv2d[l];
      fill_set(s, ch, L, 1, 1, out[0]);
      fill_set(s, ch, L, 1, 0, out[1]);
      fill_set(s, ch, L, 0, 1, out[2]);
      if (l == 1)
        return 3 | 0x100;
      ChildRec& r = out[3];
      r.id = make_node(kINodeLevel, unsigned(l - 1), 0, 0);
      r.p = gptr(ch.pyr_p)[h->pyr_nodes + l - 1];
      r.d = gptr(ch.pyr_d)[h->pyr_nodes + l - 1];
      r.kind = 2;
      r.lis_desc = unsigned(h->nlis);   // behind list 0
      r.sign = 0;
      return 4;
    }
    NodeGeom g;
    node_geom(s, nd, g);
    int k = 0;
    for (int cy = int(g.nyc) - 1; cy >= 0; cy--)
      for (int cx = int(g.nxc) - 1; cx >= 0; cx--, k++) {
        const unsigned jx = g.x0 + cx, jy = g.y0 + cy;
        ChildRec& r = out[k];
        unsigned lx, ly, lz;
        node_len(s, g.Lc, jx, jy, 0, lx, ly, lz);
        if (lx * ly == 1) {
          r.id = make_node(g.Lc, jx, jy, 0);
          r.p = node_p(s, ch, g.Lc, jx, jy, 0);
          r.kind = 0;
          r.d = 0;
          r.lis_desc = 0;
          const unsigned long long ri = node_raster(s, g.Lc, jx, jy, 0);
          r.sign = (gptr(ch.signs)[ri >> 5] >> (ri & 31)) & 1u;
        }
        else
          fill_set(s, ch, g.Lc, jx, jy, r);
      }
    return k;
  }

  static __device__ __forceinline__ int planes(const Data& t, const ChunkDev& ch, unsigned c)
  {
    return Tree3D::planes(t, ch, c);
  }

  static __device__ __forceinline__ int num_roots(const Data& t, const ChunkDev& ch, unsigned)
  {
    return t.shapes[ch.shape].h->nxf2d > 0 ? 2 : 1;
  }

  static __device__ __forceinline__ void root(const Data& t, const ChunkDev& ch, unsigned, int r,
                                              node_t& nd, unsigned& lis_desc, unsigned& order)
  {
    const ShapeHeader* h = t.shapes[ch.shape].h;
    order = 0;
    if (r == 0) {
      nd = make_node(h->lv2d[h->nxf2d], 0, 0, 0);
      lis_desc = unsigned(h->nlis - 1 - h->nxf2d);
    }
    else {
      nd = make_node(kINodeLevel, unsigned(h->nxf2d), 0, 0);
      lis_desc = unsigned(h->nlis);
    }
  }
};

// ---------------------------------------------------------------------------------------------
// host driver
// ---------------------------------------------------------------------------------------------

void Speck3DEncoder::encode(ChunkDev* d_chunks, const std::vector<ChunkDev>& h_chunks,
                            const ShapeDev* d_shapes, const std::vector<ShapeTables>& shapes,
                            std::vector<EncResult>& results, cudaStream_t st)
{
  const int nchunks = int(h_chunks.size());
  size_t max_n = 0;
  unsigned long long cap_nodes = 0;
  for (auto& c : h_chunks) {
    max_n = std::max<size_t>(max_n, c.n);
    cap_nodes += shapes[c.shape].h.set_nodes;
  }
  // pyramid, per shape group (levels of one chain are built child -> parent)
  std::vector<std::vector<int>> groups(shapes.size());
  for (int c = 0; c < nchunks; c++)
    groups[h_chunks[c].shape].push_back(c);
  ids_.reserve(nchunks * sizeof(int));
  {
    std::vector<int> flat;
    for (auto& g : groups)
      flat.insert(flat.end(), g.begin(), g.end());
    rt::h2d(ids_.p, flat.data(), flat.size() * sizeof(int), st);
    rt::sync(st);
  }
  size_t id_off = 0;
  int max_depth = 1;
  bool any_2d = false, any_3d = false;
  auto bound = [&](int c, const ChunkDev& hc) {
    const ShapeHeader& h = shapes[h_chunks[c].shape].h;
    unsigned long long bits = hc.n * (unsigned long long)(hc.planes + 1) +
                              h.set_nodes * (unsigned long long)hc.planes + 64;
    if (hc.budget != ~0ull)
      bits = std::min(bits, hc.budget + 3 * hc.n + h.set_nodes + 64);
    return bits;
  };
  // one power-of-two dyadic shape for the whole batch: shift-addressed tree (Tree3DPow2)
  if (shapes.size() == 1 && shapes[0].h.pow2 && !shapes[0].h.is2d && !std::getenv("SPERR_B200_NO_POW2ENC")) {
    const ShapeHeader& h = shapes[0].h;
    Tree3DPow2::Data tree;
    tree.shapes = d_shapes;
    Pow2Info& g = tree.g;
    g.Dx = h.ax[0].D; g.Dy = h.ax[1].D; g.Dz = h.ax[2].D;
    g.J = std::max(g.Dx, std::max(g.Dy, g.Dz));
    g.nlis = h.nlis;
    g.nx = h.nx; g.ny = h.ny;
    for (int l = 0; l < h.nlevels; l++)
      if (l != h.leaf_level && h.lv[l].chain == 0)
        g.p_off[h.lv[l].j] = h.lv[l].p_off;
    {
      rt::ProfScope ps("enc.pyramid", st);
      for (int j = g.J - 1; j >= 0; j--) {
        const size_t nodes = size_t(1) << (std::min(j, g.Dx) + std::min(j, g.Dy) + std::min(j, g.Dz));
        LAUNCH(k_pyr_pow2, dim3(unsigned((nodes + 255) / 256), unsigned(nchunks)), dim3(256), 0, st, d_chunks,
               g, j);
      }
    }
    run_encoder<Tree3DPow2>(work_, d_chunks, nchunks, max_n, tree, cap_nodes, g.J + 2, bound, results, st);
    return;
  }
  rt::ProfScope* ps_pyr = new rt::ProfScope("enc.pyramid", st);
  for (size_t si = 0; si < shapes.size(); si++) {
    const auto& g = groups[si];
    if (g.empty())
      continue;
    const ShapeHeader& h = shapes[si].h;
    const int* d_ids = ids_.as<int>() + id_off;
    id_off += g.size();
    std::vector<int> order;
    for (int l = 0; l < h.nlevels; l++)
      if (l != h.leaf_level) {
        order.push_back(l);
        max_depth = std::max(max_depth, h.lv[l].j + 2);
      }
    std::sort(order.begin(), order.end(), [&](int a, int b) { return h.lv[a].j > h.lv[b].j; });
    const int check_real = (h.dyadic < 0 && !h.is2d) ? 1 : 0;
    for (int l : order) {
      const size_t nodes = (size_t)h.lv[l].cx * h.lv[l].cy * h.lv[l].cz;
      LAUNCH(k_pyr_level, dim3(unsigned((nodes + 255) / 256), unsigned(g.size())), dim3(256), 0, st,
             d_chunks, d_shapes, d_ids, l, check_real, h.is2d);
    }
    if (h.is2d) {
      LAUNCH(k_pyr_iset, dim3(unsigned((g.size() + 63) / 64)), dim3(64), 0, st, d_chunks, d_shapes, d_ids,
             int(g.size()));
      max_depth += h.nxf2d + 1;
      any_2d = true;
    }
    else
      any_3d = true;
  }
  delete ps_pyr;
  if (any_2d && any_3d)
    throw std::runtime_error("2D and 3D chunks in one batch");
  Tree3D::Data tree{d_shapes};
  if (any_2d)
    run_encoder<Tree2D>(work_, d_chunks, nchunks, max_n, tree, cap_nodes, max_depth, bound, results, st);
  else
    run_encoder<Tree3D>(work_, d_chunks, nchunks, max_n, tree, cap_nodes, max_depth, bound, results, st);
}

}  // namespace sperr_b200
