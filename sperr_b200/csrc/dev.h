// Device-visible data structures shared by all kernels.
#pragma once

#include "geom.h"
#include "rt.h"

namespace sperr_b200 {

// Flat view of one ShapeTables object in device memory.
struct ShapeDev {
  const ShapeHeader* h;
  const uint32_t* bnd;
  const uint32_t* child0;
  const uint8_t* lev;
};

// One chunk of a batch. All pointers are device pointers into the batch's work buffers.
struct ChunkDev {
  // geometry: position inside the source volume and extent
  uint32_t x0, y0, z0;
  uint32_t nx, ny, nz;
  unsigned long long n;     // nx*ny*nz
  int shape;                // index into the batch's ShapeDev array

  double* coef;             // n fp64: conditioned values -> wavelet coefficients (in place)
  double* scratch;          // fused transform: compact approx boxes of levels 1 .. L-1
  int fused;                // dyadic shape: transformed by dwt_fused.cu (one round trip per level)
  uint32_t* obits;          // decoder: bit i set when value i has an outlier corrector
  void* mag;                // n quantised magnitudes (uint32_t or uint64_t, see `wide`)
  uint32_t* signs;          // ceil(n/32) words, bit i = 1 when value i is non-negative
  int8_t* pleaf;            // n: msb position of each magnitude (-1 for zero)
  int8_t* cmap;             // n: plane at which each coefficient's parent set turns significant
  int8_t* pyr_p;            // significance pyramid: msb of the max magnitude of every node
  uint32_t* pyr_d;          // bits a node's expansion emits in the plane it turns significant
  uint32_t* spk;            // SPECK payload staging (zero-initialised bit array)
  unsigned long long spk_cap_bits;

  // conditioner / quantiser scalars
  double mean;
  double q;
  int is_const;             // all values equal: chunk stream is the 17-byte header only
  int wide;                 // 0: uint32 magnitudes, 1: uint64
  int fe_invalid;           // quantiser saw NaN / overflow
  unsigned long long max_bits;   // bit pattern of max |coef| (fp64)
  unsigned long long min_key, max_key;   // order-preserving keys of min / max input value
  double first_val;         // v[0] as fp64 (constant test, constant header)

  // SPECK3D encoder state
  int planes;               // num_bitplanes
  int active;               // still inside the bit-plane loop
  unsigned long long cursor;      // bits produced before the current plane
  unsigned long long lis_base;    // absolute position of the current plane's LIS part
  unsigned long long budget;      // fixed-rate budget in bits (~0ull = unlimited)
  unsigned long long total_bits;  // final value of wtell()
  int last_plane;           // plane index (threshold 2^n) at which encoding stopped
  int stop_after_sort;      // 1 if the budget was met right after the sorting pass
  unsigned int ncand;       // live sets appended for the next plane
  int cur_n;                // plane being coded in this step (threshold 2^cur_n)
  int plane_live;           // chunk takes part in this step
  int next_needed;          // another plane follows: keep the lists up to date

  // outliers (PWE mode)
  unsigned int n_outliers;
};

// ---- node ids: level(8) | iz(16) | iy(16) | ix(16) ----
typedef unsigned long long node_t;
HD inline node_t make_node(int level, unsigned ix, unsigned iy, unsigned iz)
{
  return (node_t(unsigned(level)) << 48) | (node_t(iz) << 32) | (node_t(iy) << 16) | node_t(ix);
}
HD inline int node_level(node_t n) { return int(n >> 48) & 0xff; }
HD inline unsigned node_ix(node_t n) { return unsigned(n) & 0xffffu; }
HD inline unsigned node_iy(node_t n) { return unsigned(n >> 16) & 0xffffu; }
HD inline unsigned node_iz(node_t n) { return unsigned(n >> 32) & 0xffffu; }

// ---- list keys: chunk(10) | inverted LIS index(6) | bit position(46) ----
constexpr int kKeyBits = 62;
constexpr int kKeyPosBits = 46;
HD inline unsigned long long make_key(unsigned chunk, unsigned lis_desc, unsigned long long pos)
{
  return ((unsigned long long)chunk << 52) | ((unsigned long long)lis_desc << kKeyPosBits) | pos;
}
HD inline unsigned key_chunk(unsigned long long k) { return unsigned(k >> 52) & 0x3ffu; }
constexpr int kMaxBatchChunks = 1024;

// ---- table accessors ----
__device__ __forceinline__ uint32_t tab_bnd(const ShapeDev& s, int a, int d, unsigned k)
{
  return __ldg(&s.bnd[s.h->tab_off[a] + s.h->ax[a].off[d] + k]);
}
__device__ __forceinline__ uint32_t tab_child0(const ShapeDev& s, int a, int d, unsigned k)
{
  return __ldg(&s.child0[s.h->tab_off[a] + s.h->ax[a].off[d] + k]);
}
__device__ __forceinline__ uint32_t tab_lev(const ShapeDev& s, int a, int d, unsigned k)
{
  return __ldg(&s.lev[s.h->tab_off[a] + s.h->ax[a].off[d] + k]);
}

}  // namespace sperr_b200
