// Implicit set-partition tree of the 3D coder, addressed through the shape tables (geom.h):
// geometry helpers shared by the encoder and the decoder.
#pragma once

#include "speck.h"

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

__device__ __forceinline__ unsigned node_lis(const ShapeDev& s, int L, unsigned ix, unsigned iy,
                                             unsigned iz)
{
  const LevelDesc& lv = s.h->lv[L];
  return tab_lev(s, 0, lv.dx, ix) + tab_lev(s, 1, lv.dy, iy) + tab_lev(s, 2, lv.dz, iz);
}

}  // namespace sperr_b200
