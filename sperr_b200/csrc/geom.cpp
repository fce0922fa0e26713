#include "geom.h"

#include <algorithm>
#include <stdexcept>

namespace sperr_b200 {

size_t num_of_xforms(size_t len)
{
  size_t n = 0;
  while (len >= 9) {
    n++;
    len -= len / 2;
  }
  return std::min<size_t>(n, 6);
}

size_t num_of_partitions(size_t len)
{
  size_t n = 0;
  while (len > 1) {
    n++;
    len -= len / 2;
  }
  return n;
}

int can_use_dyadic(size_t nx, size_t ny, size_t nz)
{
  if (nz < 2 || ny < 2)
    return -1;
  const size_t xy = num_of_xforms(std::min(nx, ny));
  const size_t z = num_of_xforms(nz);
  if (xy == z || (xy >= 5 && z >= 5))
    return int(std::min(xy, z));
  return -1;
}

std::array<size_t, 2> calc_approx_detail_len(size_t len, size_t lev)
{
  size_t low = len, high = 0;
  for (size_t i = 0; i < lev; i++) {
    high = low / 2;
    low -= high;
  }
  return {low, high};
}

bool chunk_grid(const size_t vol[3], const size_t chunk[3], size_t nseg[3], size_t* count)
{
  size_t total = 1;
  for (int a = 0; a < 3; a++) {
    if (vol[a] == 0 || chunk[a] == 0)
      return false;
    nseg[a] = vol[a] / chunk[a];
    if (vol[a] % chunk[a] > chunk[a] / 2)
      nseg[a]++;
    if (nseg[a] == 0)
      nseg[a] = 1;
    if (total > (~size_t(0)) / nseg[a])
      return false;   // the product does not fit a size_t
    total *= nseg[a];
  }
  *count = total;
  return true;
}

bool shard_ranges(const size_t vol[3], const size_t chunk[3], size_t world, size_t* begins)
{
  size_t cd[3], g[3], total = 0;
  for (int a = 0; a < 3; a++)
    cd[a] = std::min(std::max<size_t>(1, chunk[a]), vol[a]);
  if (world == 0 || !chunk_grid(vol, cd, g, &total) || total < world)
    return false;
  // a contiguous index range [b, e) of the x-fastest chunk grid is a box iff its bounding box
  // (in chunk units) holds exactly e - b chunks
  auto is_box = [&](size_t b, size_t e) {
    size_t lo[3] = {~size_t(0), ~size_t(0), ~size_t(0)}, hi[3] = {0, 0, 0};
    // the extremes of a contiguous range are reached at its ends and at row / slab crossings
    const size_t row = g[0], slab = g[0] * g[1];
    auto visit = [&](size_t i) {
      const size_t c[3] = {i % row, (i / row) % g[1], i / slab};
      for (int k = 0; k < 3; k++) {
        lo[k] = std::min(lo[k], c[k]);
        hi[k] = std::max(hi[k], c[k]);
      }
    };
    visit(b);
    visit(e - 1);
    if (b / row != (e - 1) / row) {   // crosses a row boundary: all x are touched
      lo[0] = 0;
      hi[0] = row - 1;
    }
    if (b / slab != (e - 1) / slab) {   // crosses a slab boundary: all y are touched
      lo[1] = 0;
      hi[1] = g[1] - 1;
    }
    return (hi[0] - lo[0] + 1) * (hi[1] - lo[1] + 1) * (hi[2] - lo[2] + 1) == e - b;
  };
  const size_t units[3] = {1, g[0], g[0] * g[1]};   // chunks, whole rows, whole slabs
  for (int u = 0; u < 3; u++) {
    if (u > 0 && units[u] == units[u - 1])
      continue;
    const size_t n = total / units[u];
    if (n < world)
      break;
    const size_t base = n / world, rem = n % world;
    bool ok = true;
    size_t pos = 0;
    for (size_t r = 0; r < world && ok; r++) {
      begins[r] = pos * units[u];
      pos += base + (r < rem ? 1 : 0);
      ok = is_box(begins[r], pos * units[u]);
    }
    begins[world] = total;
    if (ok)
      return true;
  }
  return false;
}

std::vector<Chunk> chunk_volume(const size_t vol[3], const size_t chunk[3])
{
  size_t nseg[3], total = 0;
  if (!chunk_grid(vol, chunk, nseg, &total) || total > (size_t(1) << 40))
    throw std::length_error("chunk_volume: unreasonable chunk grid");
  auto seg = [&](int a, size_t i, uint32_t& beg, uint32_t& len) {
    const size_t b = i * chunk[a];
    const size_t e = (i + 1 == nseg[a]) ? vol[a] : (i + 1) * chunk[a];
    beg = uint32_t(b);
    len = uint32_t(e - b);
  };
  std::vector<Chunk> out;
  out.reserve(nseg[0] * nseg[1] * nseg[2]);
  for (size_t z = 0; z < nseg[2]; z++)
    for (size_t y = 0; y < nseg[1]; y++)
      for (size_t x = 0; x < nseg[0]; x++) {
        Chunk c;
        seg(0, x, c.x0, c.lx);
        seg(1, y, c.y0, c.ly);
        seg(2, z, c.z0, c.lz);
        out.push_back(c);
      }
  return out;
}

size_t mean_num_strides(size_t len)
{
  size_t ns = 2048;
  if (len % ns == 0)
    return ns;
  size_t num;
  for (num = ns; num <= 32768; num++)
    if (len % num == 0)
      break;
  if (len % num == 0)
    return num;
  for (num = ns; num > 0; num--)
    if (len % num == 0)
      break;
  return num;
}

namespace {

struct AxisBuild {
  std::vector<std::vector<uint32_t>> bnd, child0;
  std::vector<std::vector<uint8_t>> lev;
  int D = 0;
};

AxisBuild build_axis(uint32_t L)
{
  AxisBuild a;
  a.bnd.push_back({0, L});
  a.lev.push_back({0});
  for (;;) {
    const auto& b = a.bnd.back();
    const auto& lv = a.lev.back();
    const size_t cnt = b.size() - 1;
    bool all1 = true;
    for (size_t k = 0; k < cnt; k++)
      if (b[k + 1] - b[k] > 1)
        all1 = false;
    if (all1)
      break;
    std::vector<uint32_t> nb, c0;
    std::vector<uint8_t> nl;
    for (size_t k = 0; k < cnt; k++) {
      const uint32_t len = b[k + 1] - b[k];
      c0.push_back(uint32_t(nb.size()));
      nb.push_back(b[k]);
      nl.push_back(uint8_t(lv[k] + (len > 1 ? 1 : 0)));
      if (len > 1) {
        nb.push_back(b[k] + (len - len / 2));
        nl.push_back(uint8_t(lv[k] + 1));
      }
    }
    c0.push_back(uint32_t(nb.size()));
    nb.push_back(L);
    a.child0.push_back(c0);
    a.bnd.push_back(nb);
    a.lev.push_back(nl);
  }
  a.D = int(a.bnd.size()) - 1;
  // identity child table for the deepest depth (clamped lookups)
  std::vector<uint32_t> idc(a.bnd.back().size());
  for (size_t k = 0; k < idc.size(); k++)
    idc[k] = uint32_t(k);
  a.child0.push_back(idc);
  return a;
}

}  // namespace

ShapeTables build_shape(uint32_t nx, uint32_t ny, uint32_t nz, bool is_2d)
{
  ShapeTables t;
  ShapeHeader& h = t.h;
  h = ShapeHeader();
  h.nx = nx; h.ny = ny; h.nz = nz;
  const uint32_t dims[3] = {nx, ny, nz};
  AxisBuild ab[3];
  for (int a = 0; a < 3; a++) {
    ab[a] = build_axis(dims[a]);
    if (ab[a].D > kMaxAxisDepth - 1)
      throw std::runtime_error("chunk dimension too large");
    h.ax[a].D = ab[a].D;
    h.tab_off[a] = t.bnd.size();
    for (int d = 0; d <= ab[a].D; d++) {
      h.ax[a].cnt[d] = int(ab[a].bnd[d].size()) - 1;
      h.ax[a].off[d] = int(t.bnd.size() - h.tab_off[a]);
      for (size_t k = 0; k < ab[a].bnd[d].size(); k++) {
        t.bnd.push_back(ab[a].bnd[d][k]);
        t.child0.push_back(ab[a].child0[d][k]);
        t.lev.push_back(k < ab[a].lev[d].size() ? ab[a].lev[d][k] : 0);
      }
    }
  }
  const int D[3] = {h.ax[0].D, h.ax[1].D, h.ax[2].D};

  // ---- chains of levels ----
  h.nlevels = 0;
  auto add_level = [&](int dx, int dy, int dz, int chain, int j) {
    if (h.nlevels >= kMaxLevels)
      throw std::runtime_error("too many pyramid levels");
    LevelDesc& l = h.lv[h.nlevels];
    l.dx = dx; l.dy = dy; l.dz = dz;
    l.cx = h.ax[0].cnt[dx]; l.cy = h.ax[1].cnt[dy]; l.cz = h.ax[2].cnt[dz];
    l.chain = chain; l.j = j; l.child = -1; l.p_off = 0;
    return h.nlevels++;
  };
  h.leaf_level = add_level(D[0], D[1], D[2], -1, 0);
  std::vector<std::array<int, 3>> chain_start;   // start triple of each chain
  std::vector<int> chain_first_level;            // LevelDesc index of j = 0
  auto add_chain = [&](int sx, int sy, int sz) {
    const int c = int(chain_start.size());
    chain_start.push_back({sx, sy, sz});
    int first = -1, prev = -1;
    for (int j = 0;; j++) {
      const int dx = std::min(sx + j, D[0]), dy = std::min(sy + j, D[1]), dz = std::min(sz + j, D[2]);
      if (dx == D[0] && dy == D[1] && dz == D[2]) {
        if (prev >= 0)
          h.lv[prev].child = h.leaf_level;
        if (first < 0)
          first = h.leaf_level;
        break;
      }
      const int li = add_level(dx, dy, dz, c, j);
      if (prev >= 0)
        h.lv[prev].child = li;
      if (first < 0)
        first = li;
      prev = li;
    }
    chain_first_level.push_back(first);
    return c;
  };
  auto level_of = [&](int chain, int j) {
    // levels of a chain are contiguous, the last one may be the shared leaf grid
    int li = chain_first_level[chain];
    for (int k = 0; k < j; k++)
      li = h.lv[li].child;
    return li;
  };
  add_chain(0, 0, 0);

  if (is_2d) {
    // ---- SPECK2D_INT::m_initialize_lists (src/SPECK2D_INT.cpp:187-213) ----
    if (nz != 1)
      throw std::runtime_error("2D shape with more than one plane");
    h.is2d = 1;
    h.dyadic = -1;
    h.nxf2d = int(num_of_xforms(std::min(nx, ny)));
    h.nlis = int(num_of_partitions(std::max(nx, ny)) + 1);
    if (h.nlis > kMaxLis - 1)
      throw std::runtime_error("too many LIS lists");
    for (int j = 0; j < kMaxAxisDepth + 2; j++)
      h.lv2d[j] = -1;
    {
      int li = chain_first_level[0];
      for (int j = 0; li >= 0 && j < kMaxAxisDepth + 2; j++) {
        h.lv2d[j] = li;
        if (li == h.leaf_level)
          break;
        li = h.lv[li].child;
      }
    }
    h.ngroups = 0;
    h.nroots = 1;
    RootDesc r;
    r.level = h.lv2d[h.nxf2d];
    r.ix = r.iy = r.iz = 0;
    r.lis = h.nxf2d;
    r.order = 0;
    h.roots[0] = r;
    unsigned long long off = 0;
    for (int i = 0; i < h.nlevels; i++) {
      if (i == h.leaf_level)
        continue;
      h.lv[i].p_off = off;
      off += (unsigned long long)h.lv[i].cx * h.lv[i].cy * h.lv[i].cz;
    }
    h.pyr_nodes = off;
    h.set_nodes = off + 8;   // + the I sets
    // list l holds sets of part_level l = nodes of chain position l
    h.lis_off[0] = 0;
    for (int l = 0; l < kMaxLis; l++) {
      unsigned long long cap = 0;
      if (l < kMaxAxisDepth + 2 && h.lv2d[l] >= 0 && h.lv2d[l] != h.leaf_level)
        cap = (unsigned long long)h.lv[h.lv2d[l]].cx * h.lv[h.lv2d[l]].cy;
      h.lis_off[l + 1] = h.lis_off[l] + cap;
    }
    h.pow2 = 0;
    return t;
  }

  // ---- initial LIS (m_initialize_lists) ----
  h.dyadic = can_use_dyadic(nx, ny, nz);
  size_t nxy, nzz;
  if (h.dyadic >= 0)
    nxy = nzz = size_t(h.dyadic);
  else {
    nxy = num_of_xforms(std::min(nx, ny));
    nzz = num_of_xforms(nz);
  }
  h.nlis = int(num_of_partitions(nx) + num_of_partitions(ny) + num_of_partitions(nz) + 1);
  std::vector<std::vector<RootDesc>> lists(h.nlis);
  h.ngroups = 0;
  int bchain = 0, bj = 0;           // where `big` lives
  uint32_t blen[3] = {nx, ny, nz};  // extent of `big` (always anchored at the origin)
  auto lis_of = [&](int level, int ix, int iy, int iz) {
    const LevelDesc& l = h.lv[level];
    return int(ab[0].lev[l.dx][ix]) + int(ab[1].lev[l.dy][iy]) + int(ab[2].lev[l.dz][iz]);
  };
  auto add_group = [&](int chain, int j, const uint32_t outer[3], const uint32_t inner[3]) {
    if (h.ngroups >= kMaxGroups)
      throw std::runtime_error("too many root groups");
    GroupDesc& g = h.grp[h.ngroups++];
    g.chain = chain; g.j_root = j;
    g.ox = int(outer[0]); g.oy = int(outer[1]); g.oz = int(outer[2]);
    g.ix = int(inner[0]); g.iy = int(inner[1]); g.iz = int(inner[2]);
  };
  size_t xf = 0;
  auto split = [&](bool sx, bool sy, bool sz) {
    // children of big = interval 0 on every axis; which axes advance is given by the flags
    const int pl = level_of(bchain, bj);
    const LevelDesc P = h.lv[pl];
    int cchain, cj;
    if (sx && sy && sz) {
      cchain = bchain;
      cj = bj + 1;
    }
    else {
      cchain = add_chain(std::min(P.dx + (sx ? 1 : 0), D[0]), std::min(P.dy + (sy ? 1 : 0), D[1]),
                         std::min(P.dz + (sz ? 1 : 0), D[2]));
      cj = 0;
    }
    const int cl = level_of(cchain, cj);
    const LevelDesc C = h.lv[cl];
    // number of children per axis: 2 if that axis advanced and the interval really split
    const int n0 = (C.dx != P.dx && blen[0] > 1) ? 2 : 1;
    const int n1 = (C.dy != P.dy && blen[1] > 1) ? 2 : 1;
    const int n2 = (C.dz != P.dz && blen[2] > 1) ? 2 : 1;
    for (int cz = 0; cz < n2; cz++)
      for (int cy = 0; cy < n1; cy++)
        for (int cx = 0; cx < n0; cx++) {
          if (cx == 0 && cy == 0 && cz == 0)
            continue;
          RootDesc r;
          r.level = cl; r.ix = cx; r.iy = cy; r.iz = cz;
          r.lis = lis_of(cl, cx, cy, cz);
          r.order = 0;
          lists[r.lis].push_back(r);
        }
    uint32_t inner[3] = {blen[0], blen[1], blen[2]};
    if (n0 == 2) inner[0] = blen[0] - blen[0] / 2;
    if (n1 == 2) inner[1] = blen[1] - blen[1] / 2;
    if (n2 == 2) inner[2] = blen[2] - blen[2] / 2;
    add_group(cchain, cj, blen, inner);
    blen[0] = inner[0]; blen[1] = inner[1]; blen[2] = inner[2];
    bchain = cchain;
    bj = cj;
  };
  while (xf < nxy && xf < nzz) {
    split(true, true, true);
    xf++;
  }
  while (xf < nxy) {
    split(true, true, false);
    xf++;
  }
  while (xf < nzz) {
    split(false, false, true);
    xf++;
  }
  {
    RootDesc r;
    r.level = level_of(bchain, bj);
    r.ix = r.iy = r.iz = 0;
    r.lis = lis_of(r.level, 0, 0, 0);
    r.order = 0;
    lists[r.lis].insert(lists[r.lis].begin(), r);
    const uint32_t none[3] = {0, 0, 0};
    add_group(bchain, bj, blen, none);
  }
  h.nroots = 0;
  for (int l = h.nlis - 1; l >= 0; l--)
    for (size_t k = 0; k < lists[l].size(); k++) {
      if (h.nroots >= kMaxRoots)
        throw std::runtime_error("too many initial sets");
      RootDesc r = lists[l][k];
      r.order = int(k);
      h.roots[h.nroots++] = r;
    }

  // ---- pyramid layout ----
  unsigned long long off = 0;
  for (int i = 0; i < h.nlevels; i++) {
    if (i == h.leaf_level)
      continue;
    h.lv[i].p_off = off;
    off += (unsigned long long)h.lv[i].cx * h.lv[i].cy * h.lv[i].cz;
  }
  h.pyr_nodes = off;
  h.set_nodes = off;

  // ---- per-list capacity: histogram of the LIS index over all pyramid nodes ----
  if (h.nlis > kMaxLis)
    throw std::runtime_error("too many LIS lists");
  std::vector<unsigned long long> cap(kMaxLis, 0);
  for (int i = 0; i < h.nlevels; i++) {
    if (i == h.leaf_level)
      continue;
    const LevelDesc& l = h.lv[i];
    const int dd[3] = {l.dx, l.dy, l.dz};
    std::vector<unsigned long long> hist[3];
    for (int a = 0; a < 3; a++) {
      hist[a].assign(kMaxAxisDepth + 2, 0);
      for (uint8_t v : ab[a].lev[dd[a]])
        hist[a][v]++;
    }
    for (int x = 0; x <= kMaxAxisDepth; x++)
      for (int y = 0; y <= kMaxAxisDepth && hist[0][x]; y++)
        for (int z = 0; z <= kMaxAxisDepth && hist[1][y]; z++)
          if (hist[2][z] && x + y + z < kMaxLis)
            cap[x + y + z] += hist[0][x] * hist[1][y] * hist[2][z];
  }
  for (int r = 0; r < h.nroots; r++)
    cap[h.roots[r].lis] += 1;  // the initial sets (the root of chain 0 is not a pyramid child)
  h.lis_off[0] = 0;
  for (int l = 0; l < kMaxLis; l++)
    h.lis_off[l + 1] = h.lis_off[l] + cap[l];
  auto is_pow2 = [](uint32_t v) { return v != 0 && (v & (v - 1)) == 0; };
  h.pow2 = (h.dyadic >= 0 && chain_start.size() == 1 && is_pow2(nx) && is_pow2(ny) && is_pow2(nz) &&
            h.leaf_level == 0)
               ? 1
               : 0;
  return t;
}

}  // namespace sperr_b200
