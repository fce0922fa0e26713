// Host-side geometry: chunking, transform-level rules and the set-partition tables the SPECK
// kernels index with. Restates (does not copy) the reference helpers:
//   num_of_xforms            /root/reference/src/sperr_helper.cpp:36-49
//   can_use_dyadic           /root/reference/src/sperr_helper.cpp:51-68
//   num_of_partitions        /root/reference/src/sperr_helper.cpp:125-134
//   calc_approx_detail_len   /root/reference/src/sperr_helper.cpp:136-146
//   chunk_volume             /root/reference/src/sperr_helper.cpp:542-592
//   SPECK3D_INT::m_initialize_lists / m_partition_S_*  /root/reference/src/SPECK3D_INT.cpp:22-97,214-430
#pragma once

#include <array>
#include <cstddef>
#include <cstdint>
#include <vector>

namespace sperr_b200 {

size_t num_of_xforms(size_t len);
size_t num_of_partitions(size_t len);
int can_use_dyadic(size_t nx, size_t ny, size_t nz);  // number of levels, or -1
std::array<size_t, 2> calc_approx_detail_len(size_t len, size_t lev);

struct Chunk {
  uint32_t x0, lx, y0, ly, z0, lz;
  size_t nelem() const { return size_t(lx) * ly * lz; }
};
std::vector<Chunk> chunk_volume(const size_t vol[3], const size_t chunk[3]);
// Segments per axis and their product (false when an extent is zero or the product overflows):
// what callers that hold UNTRUSTED header dimensions check before chunk_volume builds its vector.
bool chunk_grid(const size_t vol[3], const size_t chunk[3], size_t nseg[3], size_t* count);
// Partition of chunk_volume's order over `world` ranks such that every rank's contiguous range
// [begins[r], begins[r + 1]) is exactly a box of chunks (so the bounding box a rank holds contains
// no chunk of another rank). Tries an even split of chunks, then of whole rows, then of whole
// z-slabs; false when none of them gives boxes or there are fewer chunks than ranks.
bool shard_ranges(const size_t vol[3], const size_t chunk[3], size_t world, size_t* begins);

// Number of strides the conditioner's mean uses (Conditioner::m_adjust_strides,
// /root/reference/src/Conditioner.cpp:137-163).
size_t mean_num_strides(size_t len);

// ---------------------------------------------------------------------------------------------
// SPECK3D set-partition tables.
//
// Every set the coder can ever create is a box whose extent along each axis is an interval of that
// axis' recursive "first half gets the ceiling" bisection. So a set is (per-axis depth, per-axis
// interval index). All XYZ splits advance the three depths together (clamped at the depth where an
// axis has reached single samples), which lets us lay the significance pyramid out as a short
// chain of dense 3D arrays ("levels"). Wavelet-packet chunks (XY-only or Z-only initial splits)
// need a few extra chains whose xy and z depths are offset.
// ---------------------------------------------------------------------------------------------

constexpr int kMaxAxisDepth = 17;  // dims <= 65535
constexpr int kMaxLevels = 96;
constexpr int kMaxRoots = 64;
constexpr int kMaxGroups = 16;
constexpr int kMaxLis = 64;

struct AxisTab {
  int D;                        // deepest depth (all intervals have length 1)
  int cnt[kMaxAxisDepth + 1];   // number of intervals at depth d
  int off[kMaxAxisDepth + 1];   // offset of depth d inside bnd / child0 / lev (cnt+1 entries each)
};

struct LevelDesc {
  int dx, dy, dz;     // per-axis depths (already clamped)
  int cx, cy, cz;     // node counts per axis
  int child;          // index of the level holding this level's children; -1 for the leaf grid
  int chain;          // which chain this level belongs to
  int j;              // position inside the chain
  unsigned long long p_off;  // offset of this level in the per-chunk pyramid arrays
};

struct RootDesc {
  int level;            // LevelDesc index
  int ix, iy, iz;       // interval indices
  int lis;              // LIS list index (SPECK "level")
  int order;            // insertion position inside its list
};

struct GroupDesc {   // region of one initial split step: outer box minus inner box (both at origin)
  int chain, j_root;
  int ox, oy, oz, ix, iy, iz;
};

// POD that is copied to the device as-is; the three axis arrays follow it in one allocation.
struct ShapeHeader {
  uint32_t nx, ny, nz;
  int dyadic;            // levels, or -1 for wavelet-packet
  AxisTab ax[3];
  unsigned long long tab_off[3];  // offset (in elements) of each axis inside the flat tables
  int nlevels;
  int leaf_level;
  LevelDesc lv[kMaxLevels];
  int nroots;
  RootDesc roots[kMaxRoots];
  int ngroups;
  GroupDesc grp[kMaxGroups];
  int nlis;                       // number of LIS lists
  unsigned long long pyr_nodes;   // total nodes over all non-leaf levels (pyramid array length)
  unsigned long long set_nodes;   // upper bound on the number of sets (nodes with > 1 element)
  // decoder list storage: list l may hold up to lis_off[l + 1] - lis_off[l] sets (every node of
  // the pyramid whose LIS index is l)
  unsigned long long lis_off[kMaxLis + 1];
  // 1 when every extent is a power of two and the chunk is dyadic: all sets are then aligned
  // boxes of a single chain and the coders can address them with shifts instead of the tables
  int pow2;
  // 2D slices (SPECK2D_INT, /root/reference/src/SPECK2D_INT.cpp:10-218): quadtree sets of one chain
  // plus the set I = everything outside the transform's approximation band. nxf2d = number of
  // transform levels = initial part_level of I; lv2d[j] = LevelDesc index of chain position j
  // (= part_level j; the S sets cut out of I at part_level l are nodes (1,1), (1,0), (0,1) of it).
  int is2d;
  int nxf2d;
  int lv2d[kMaxAxisDepth + 2];
};

struct ShapeTables {
  ShapeHeader h;
  std::vector<uint32_t> bnd;     // interval boundaries
  std::vector<uint32_t> child0;  // index at depth d+1 of the first child
  std::vector<uint8_t> lev;      // number of real splits from the root interval
};

ShapeTables build_shape(uint32_t nx, uint32_t ny, uint32_t nz, bool is_2d = false);

}  // namespace sperr_b200
