// Fused 3D CDF 9/7 lifting: one HBM round trip per transform level.
//
// Reference behaviour reproduced bit-exactly (STRICT flavour, see transform.cu):
//   CDF97::m_dwt3d_dyadic / m_idwt3d_dyadic        /root/reference/src/CDF97.cpp:284-302
//   m_dwt3d_one_level / m_idwt3d_one_level         /root/reference/src/CDF97.cpp:387-474
//   QccWAVCDF97AnalysisSymmetric / Synthesis       /root/reference/src/CDF97.cpp:598-666
// and, fused into level 0, Conditioner's mean subtraction / addition (src/Conditioner.cpp:46-64,
// 82-96), m_gather_chunk / m_scatter_chunk (src/SPERR3D_OMP_C.cpp:237-261, SPERR3D_OMP_D.cpp:
// 167-184), the outlier scan of PWE mode (src/SPECK_FLT.cpp:468-474) and the coefficient maximum
// the quantiser needs (src/SPECK_FLT.cpp:318-321).
//
// How (not how the reference does it). The reference transforms all X lines, then all Y lines, then
// all Z lines of a level, each pass a full sweep over memory. A lifting step only looks one sample
// to each side, so after the four steps an output depends on inputs at most 4 samples away: a tile
// with a 4-sample halo yields exactly the reference's values for its interior, and the symmetric
// boundary rule of the reference is the same as loading the halo through whole-sample mirroring
// (x[-k] = x[k], x[L-1+k] = x[L-1-k]; floating-point addition commutes, so every lifting step
// preserves the symmetry bit for bit). A CTA owns a 32 x 32 (x, y) tile and marches along z one
// sample pair at a time:
//   forward: load two (40 x 40) planes -> lift the rows in shared memory (one thread per row,
//            streaming, in place) -> lift the columns -> every thread feeds its (x, y) columns into
//            a streaming lifting state held in registers -> write de-interleaved coefficients;
//   inverse: the mirror image (z in registers first, then columns, then rows).
// Levels ping-pong through compact per-level boxes (ChunkDev::scratch) so that no CTA reads what
// another one is overwriting; detail sub-bands are written once to their final place in `coef`.
#include "kernels.h"

namespace sperr_b200 {

constexpr int kFT = 32;          // tile interior (samples per axis)
constexpr int kFH = 4;           // halo
constexpr int kFI = kFT + 2 * kFH;   // 40 samples loaded per axis = 20 pairs
constexpr int kFP = kFI + 1;     // shared-memory pitch (doubles)
constexpr int kFNP = kFI / 2;    // pairs per line
constexpr int kFThreads = 256;

__device__ __forceinline__ int mirror(int i, int L)
{
  if (i < 0)
    i = -i;
  if (i >= L)
    i = 2 * (L - 1) - i;
  return i < 0 ? 0 : (i >= L ? L - 1 : i);   // further out than one reflection: value never used
}

// ---- streaming lifting: feed sample pairs in order, get the pair two positions back ----

struct FwdState {
  double E, O, O1, E1, O2;
};
__device__ __forceinline__ void fwd_step(const CdfC& k, FwdState& s, double e, double o, double& e2,
                                         double& o3)
{
  const double o1n = __dadd_rn(s.O, __dmul_rn(k.ALPHA, __dadd_rn(s.E, e)));
  const double e1n = __dadd_rn(s.E, __dmul_rn(k.BETA, __dadd_rn(s.O1, o1n)));
  const double o2n = __dadd_rn(s.O1, __dmul_rn(k.GAMMA, __dadd_rn(s.E1, e1n)));
  e2 = __dmul_rn(k.EPSILON, __dadd_rn(s.E1, __dmul_rn(k.DELTA, __dadd_rn(s.O2, o2n))));
  o3 = __dmul_rn(o2n, -k.INV_EPSILON);
  s.E = e; s.O = o; s.O1 = o1n; s.E1 = e1n; s.O2 = o2n;
}

struct InvState {
  double OP, E1, O1, E2;
};
__device__ __forceinline__ void inv_step(const CdfC& k, InvState& s, double e, double o, double& x0,
                                         double& x1)
{
  const double opn = __dmul_rn(o, -k.EPSILON);
  const double e1n = __dsub_rn(__dmul_rn(e, k.INV_EPSILON), __dmul_rn(k.DELTA, __dadd_rn(s.OP, opn)));
  const double o1n = __dsub_rn(s.OP, __dmul_rn(k.GAMMA, __dadd_rn(s.E1, e1n)));
  const double e2n = __dsub_rn(s.E1, __dmul_rn(k.BETA, __dadd_rn(s.O1, o1n)));
  x0 = s.E2;
  x1 = __dsub_rn(s.O1, __dmul_rn(k.ALPHA, __dadd_rn(s.E2, e2n)));
  s.OP = opn; s.E1 = e1n; s.O1 = o1n; s.E2 = e2n;
}

// In-place lifting of one line of kFNP pairs in shared memory (element i at p[i * stride]); on
// return pairs 2 .. kFNP-3 hold the transformed values (forward: e, o interleaved as they came;
// inverse: reconstructed samples), the outer ones are scratch.
template <bool INVERSE>
__device__ __forceinline__ void lift_line(const CdfC& k, double* p, int stride)
{
  if (!INVERSE) {
    FwdState s = {0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int j = 0; j < kFNP; j++) {
      double a, b;
      fwd_step(k, s, p[(2 * j) * stride], p[(2 * j + 1) * stride], a, b);
      if (j >= 4) {
        p[(2 * j - 4) * stride] = a;
        p[(2 * j - 3) * stride] = b;
      }
    }
  }
  else {
    InvState s = {0.0, 0.0, 0.0, 0.0};
#pragma unroll 4
    for (int j = 0; j < kFNP; j++) {
      double a, b;
      inv_step(k, s, p[(2 * j) * stride], p[(2 * j + 1) * stride], a, b);
      if (j >= 4) {
        p[(2 * j - 4) * stride] = a;
        p[(2 * j - 3) * stride] = b;
      }
    }
  }
}

struct FusedArgs {
  const ChunkDev* chunks;
  const int* ids;
  SrcVol vol;           // level 0: source volume (forward) / destination volume (inverse, mode 1)
  int lx, ly, lz;       // box of this level
  long long src_off;    // forward, level > 0: the box sits compact at scratch + src_off
  long long apx_off;    // approx box of this level: compact at scratch + apx_off, or -1: in coef
  long long out_off;    // inverse, mode 0: the rebuilt box goes compact to scratch + out_off, or -1: coef
  int tiles_x, tiles_y, zsegs;
  int last;             // forward: this is the coarsest level (its approx band is final)
  double tol;           // inverse mode 2
  OutlierSink sink;     // inverse mode 2: where outliers are recorded
  CorrectorList cor;    // inverse mode 1: outlier correctors to add before the mean (may be empty)
  CdfC k;
};

// SRC 0: float volume, 1: double volume (both minus the chunk mean), 2: compact fp64 box in scratch
template <int SRC>
__global__ void __launch_bounds__(kFThreads, 3) k_fwd3d(FusedArgs a)
{
  __shared__ double tile[2][kFI][kFP];
  __shared__ unsigned long long s_max;
  const ChunkDev& ch = a.chunks[a.ids[blockIdx.y]];
  if (ch.is_const)
    return;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  int bb = blockIdx.x;
  const int X0 = (bb % a.tiles_x) * kFT;
  bb /= a.tiles_x;
  const int Y0 = (bb % a.tiles_y) * kFT;
  const int seg = bb / a.tiles_y;
  const int lx = a.lx, ly = a.ly, lz = a.lz;
  const int ax = lx - lx / 2, ay = ly - ly / 2, az = lz - lz / 2;
  const int pps = (az + a.zsegs - 1) / a.zsegs;
  const int k0 = seg * pps, k1 = min(az, k0 + pps);
  if (k0 >= k1)
    return;
  if (tid == 0)
    s_max = 0;
  const CdfC k = a.k;
  const double mean = ch.mean;

  // what this thread loads of every plane: tile elements tid, tid + 256, ...
  constexpr int kPer = (kFI * kFI + kFThreads - 1) / kFThreads;   // 7
  unsigned long long goff[kPer];
  unsigned short sidx[kPer];
  for (int s = 0; s < kPer; s++) {
    const int idx = tid + s * kFThreads;
    const int ty = idx / kFI, tx = idx % kFI;
    sidx[s] = (unsigned short)(ty * kFP + tx);
    const int gx = mirror(X0 - kFH + tx, lx), gy = mirror(Y0 - kFH + ty, ly);
    if (SRC == 2)
      goff[s] = (unsigned long long)gy * lx + gx;
    else
      goff[s] = (unsigned long long)(ch.y0 + gy) * a.vol.vx + (ch.x0 + gx);
  }
  const unsigned long long plane = SRC == 2 ? (unsigned long long)lx * ly : a.vol.vx * a.vol.vy;
  const double* sbox = SRC == 2 ? ch.scratch + a.src_off : nullptr;

  FwdState st[4];
  for (int c = 0; c < 4; c++)
    st[c] = FwdState{0.0, 0.0, 0.0, 0.0, 0.0};
  unsigned long long vmax = 0;
  const size_t cnx = ch.nx, cnxy = (size_t)ch.nx * ch.ny;

  for (int j = k0 - 2; j <= k1 + 1; j++) {
    // ---- load planes 2j and 2j + 1 (mirrored) ----
    for (int p = 0; p < 2; p++) {
      const int gz = mirror(2 * j + p, lz);
      const unsigned long long zoff = (SRC == 2 ? (unsigned long long)gz : (unsigned long long)(ch.z0 + gz)) * plane;
      for (int s = 0; s < kPer; s++) {
        const int idx = tid + s * kFThreads;
        if (idx < kFI * kFI) {
          double v;
          if (SRC == 0)
            v = __dsub_rn(double(reinterpret_cast<const float*>(a.vol.ptr)[zoff + goff[s]]), mean);
          else if (SRC == 1)
            v = __dsub_rn(reinterpret_cast<const double*>(a.vol.ptr)[zoff + goff[s]], mean);
          else
            v = sbox[zoff + goff[s]];
          (&tile[p][0][0])[sidx[s]] = v;
        }
      }
    }
    __syncthreads();
    // ---- rows (x) ----
    if (tid < 2 * kFI)
      lift_line<false>(k, &tile[tid / kFI][tid % kFI][0], 1);
    __syncthreads();
    // ---- columns (y), only the x positions that are valid after the row pass ----
    if (tid < 2 * kFT)
      lift_line<false>(k, &tile[tid / kFT][0][kFH + tid % kFT], kFP);
    __syncthreads();
    // ---- z: streaming state in registers, 4 (x, y) columns per thread ----
    const int kk = j - 2;   // output pair index
    const bool emit = kk >= k0 && kk < k1;
    const int x = X0 + lane;
    const int px = x & 1;
    const int xo = (x >> 1) + (px ? ax : 0);
    for (int c = 0; c < 4; c++) {
      const int ry = warp + 8 * c;
      double e2, o3;
      fwd_step(k, st[c], tile[0][kFH + ry][kFH + lane], tile[1][kFH + ry][kFH + lane], e2, o3);
      const int y = Y0 + ry;
      if (emit && x < lx && y < ly) {
        const int py = y & 1;
        const int yo = (y >> 1) + (py ? ay : 0);
        // even z output (plane kk of the low band)
        if (px == 0 && py == 0) {   // approx band of this level
          if (a.apx_off >= 0)
            ch.scratch[a.apx_off + ((size_t)kk * ay + (y >> 1)) * ax + (x >> 1)] = e2;
          else
            ch.coef[(size_t)kk * cnxy + (size_t)(y >> 1) * cnx + (x >> 1)] = e2;
          if (a.last) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(e2) & 0x7fffffffffffffffull;
            vmax = b > vmax ? b : vmax;
          }
        }
        else {
          ch.coef[(size_t)kk * cnxy + (size_t)yo * cnx + xo] = e2;
          const unsigned long long b = (unsigned long long)__double_as_longlong(e2) & 0x7fffffffffffffffull;
          vmax = b > vmax ? b : vmax;
        }
        if (kk < lz / 2) {   // odd z output (plane az + kk)
          ch.coef[(size_t)(az + kk) * cnxy + (size_t)yo * cnx + xo] = o3;
          const unsigned long long b = (unsigned long long)__double_as_longlong(o3) & 0x7fffffffffffffffull;
          vmax = b > vmax ? b : vmax;
        }
      }
    }
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t = __shfl_xor_sync(0xffffffffu, vmax, o);
    vmax = t > vmax ? t : vmax;
  }
  if (lane == 0 && vmax)
    atomicMax(&s_max, vmax);
  __syncthreads();
  if (tid == 0 && s_max)
    atomicMax(const_cast<unsigned long long*>(&ch.max_bits), s_max);
}

// OUT 0: fp64 box in scratch (levels > 0) or, at level 0, raw fp64 values into the volume `vol`,
//     1: destination volume (+ outlier corrector, + mean, float or double),
//     2: compare with the source volume and record the outliers
template <int OUT>
__global__ void __launch_bounds__(kFThreads, 2) k_inv3d(FusedArgs a)
{
  __shared__ double tile[2][kFI][kFP];
  const ChunkDev& ch = a.chunks[a.ids[blockIdx.y]];
  if (ch.is_const)
    return;
  const int tid = threadIdx.x;
  int bb = blockIdx.x;
  const int X0 = (bb % a.tiles_x) * kFT;
  bb /= a.tiles_x;
  const int Y0 = (bb % a.tiles_y) * kFT;
  const int seg = bb / a.tiles_y;
  const int lx = a.lx, ly = a.ly, lz = a.lz;
  const int ax = lx - lx / 2, ay = ly - ly / 2, az = lz - lz / 2;
  const int pps = (az + a.zsegs - 1) / a.zsegs;
  const int k0 = seg * pps, k1 = min(az, k0 + pps);
  if (k0 >= k1)
    return;
  const CdfC k = a.k;
  const size_t cnx = ch.nx, cnxy = (size_t)ch.nx * ch.ny;

  // the (x, y) columns this thread runs the z lifting for: tile elements tid, tid + 256, ...
  constexpr int kPer = (kFI * kFI + kFThreads - 1) / kFThreads;   // 7
  unsigned long long coff[kPer];   // offset inside a z plane of coef
  long long aoff[kPer];            // offset inside a z plane of the approx box, or -1: not approx in (x, y)
  InvState st[kPer];
  unsigned short sidx[kPer];
  for (int s = 0; s < kPer; s++) {
    st[s] = InvState{0.0, 0.0, 0.0, 0.0};
    const int idx = tid + s * kFThreads;
    const int ty = idx / kFI, tx = idx % kFI;
    sidx[s] = (unsigned short)(ty * kFP + tx);
    const int gx = mirror(X0 - kFH + tx, lx), gy = mirror(Y0 - kFH + ty, ly);
    const int xo = (gx >> 1) + ((gx & 1) ? ax : 0), yo = (gy >> 1) + ((gy & 1) ? ay : 0);
    coff[s] = (unsigned long long)yo * cnx + xo;
    aoff[s] = ((gx | gy) & 1) ? -1ll
                              : (a.apx_off >= 0 ? (long long)(gy >> 1) * ax + (gx >> 1)
                                                : (long long)((size_t)(gy >> 1) * cnx + (gx >> 1)));
  }
  const double* abox = a.apx_off >= 0 ? ch.scratch + a.apx_off : ch.coef;
  const size_t aplane = a.apx_off >= 0 ? (size_t)ax * ay : cnxy;

  for (int j = k0 - 2; j <= k1 + 1; j++) {
    // ---- z: pair j = (low-band plane mirror(2j) / 2, high-band plane az + mirror(2j + 1) / 2) ----
    const int ze = mirror(2 * j, lz) >> 1, zo = az + (mirror(2 * j + 1, lz) >> 1);
    for (int s = 0; s < kPer; s++) {
      const int idx = tid + s * kFThreads;
      if (idx < kFI * kFI) {
        const double e = aoff[s] >= 0 ? abox[(size_t)ze * aplane + aoff[s]] : ch.coef[(size_t)ze * cnxy + coff[s]];
        const double o = ch.coef[(size_t)zo * cnxy + coff[s]];
        double x0, x1;
        inv_step(k, st[s], e, o, x0, x1);
        (&tile[0][0][0])[sidx[s]] = x0;
        (&tile[1][0][0])[sidx[s]] = x1;
      }
    }
    __syncthreads();
    const int kk = j - 2;   // planes 2 kk and 2 kk + 1 of the rebuilt box
    if (kk >= k0 && kk < k1) {   // block-uniform
      // ---- columns (y) over all x of the tile, then rows (x) over the valid y ----
      if (tid < 2 * kFI)
        lift_line<true>(k, &tile[tid / kFI][0][tid % kFI], kFP);
      __syncthreads();
      if (tid < 2 * kFT)
        lift_line<true>(k, &tile[tid / kFT][kFH + tid % kFT][0], 1);
      __syncthreads();
      // ---- epilogue: 2 planes x 32 x 32 values, 8 per thread ----
      for (int s = 0; s < 8; s++) {
        const int idx = tid + s * kFThreads;
        const int p = idx >> 10, ry = (idx >> 5) & 31, rx = idx & 31;
        const int x = X0 + rx, y = Y0 + ry, z = 2 * kk + p;
        if (x < lx && y < ly && z < lz) {
          const double v = tile[p][kFH + ry][kFH + rx];
          if (OUT == 0 && a.out_off >= 0) {
            ch.scratch[a.out_off + ((size_t)z * ly + y) * lx + x] = v;
          }
          else {
            const unsigned long long g = (unsigned long long)(ch.z0 + z) * a.vol.vx * a.vol.vy +
                                         (unsigned long long)(ch.y0 + y) * a.vol.vx + (ch.x0 + x);
            if (OUT == 0) {
              reinterpret_cast<double*>(const_cast<void*>(a.vol.ptr))[g] = v;
            }
            else if (OUT == 1) {
              double w = v;
              if (a.cor.key) {   // src/SPECK_FLT.cpp:576-585: correctors are added before the mean
                const unsigned long long i = (unsigned long long)z * cnxy + (size_t)y * cnx + x;
                if ((ch.obits[i >> 5] >> (i & 31)) & 1u)
                  w = __dadd_rn(w, corrector_lookup(a.cor, a.ids[blockIdx.y], i));
              }
              w = __dadd_rn(w, ch.mean);
              if (a.vol.is_float)
                reinterpret_cast<float*>(const_cast<void*>(a.vol.ptr))[g] = __double2float_rn(w);
              else
                reinterpret_cast<double*>(const_cast<void*>(a.vol.ptr))[g] = w;
            }
            else {
              const double orig = a.vol.is_float ? double(reinterpret_cast<const float*>(a.vol.ptr)[g])
                                                 : reinterpret_cast<const double*>(a.vol.ptr)[g];
              const double diff = __dsub_rn(__dsub_rn(orig, ch.mean), v);
              if (fabs(diff) > a.tol)
                outlier_append(a.sink, a.ids[blockIdx.y], (unsigned long long)z * cnxy + (size_t)y * cnx + x,
                               diff);
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// Doubles of ChunkDev::scratch a chunk of this shape needs (compact boxes of levels 1 .. L-1), and
// the offset of every level's box.
size_t fused_scratch_elems(uint32_t nx, uint32_t ny, uint32_t nz, long long off[8])
{
  const int L = can_use_dyadic(nx, ny, nz);
  size_t tot = 0;
  for (int l = 0; l < 8; l++)
    off[l] = -1;
  for (int l = 1; l < L; l++) {
    off[l] = (long long)tot;
    tot += calc_approx_detail_len(nx, l)[0] * calc_approx_detail_len(ny, l)[0] *
           calc_approx_detail_len(nz, l)[0];
  }
  return tot;
}

static void fused_grid(FusedArgs& a, int nids, dim3& grid)
{
  a.tiles_x = (a.lx + kFT - 1) / kFT;
  a.tiles_y = (a.ly + kFT - 1) / kFT;
  const int az = a.lz - a.lz / 2;
  // split z when the grid would not fill the GPU a few times over (each segment re-reads 4 pairs)
  int zs = 1;
  while ((long long)a.tiles_x * a.tiles_y * nids * zs < 148 * 3 * 4 && az / (zs * 2) >= 16)
    zs *= 2;
  a.zsegs = zs;
  grid = dim3(unsigned(a.tiles_x * a.tiles_y * zs), unsigned(nids));
}

CdfC cdf_constants();

// Forward transform of dyadic chunks straight from the source volume: coef receives the final
// coefficients, ChunkDev::max_bits the largest magnitude.
void launch_dwt_fused_forward(const SrcVol& src, const ChunkDev* d_chunks, const int* d_ids, int nids,
                              uint32_t nx, uint32_t ny, uint32_t nz, cudaStream_t st)
{
  const int L = can_use_dyadic(nx, ny, nz);
  long long off[8];
  fused_scratch_elems(nx, ny, nz, off);
  FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  a.chunks = d_chunks;
  a.ids = d_ids;
  a.vol = src;
  a.k = cdf_constants();
  for (int l = 0; l < L; l++) {
    a.lx = int(calc_approx_detail_len(nx, l)[0]);
    a.ly = int(calc_approx_detail_len(ny, l)[0]);
    a.lz = int(calc_approx_detail_len(nz, l)[0]);
    a.src_off = l == 0 ? -1 : off[l];
    a.apx_off = l + 1 < L ? off[l + 1] : -1;
    a.last = l + 1 == L;
    dim3 grid;
    fused_grid(a, nids, grid);
    if (l > 0)
      LAUNCH(k_fwd3d<2>, grid, dim3(kFThreads), 0, st, a);
    else if (src.is_float)
      LAUNCH(k_fwd3d<0>, grid, dim3(kFThreads), 0, st, a);
    else
      LAUNCH(k_fwd3d<1>, grid, dim3(kFThreads), 0, st, a);
  }
}

// Inverse transform of dyadic chunks from coef. mode 0: the rebuilt fp64 values are written into
// the (double) volume `vol`; mode 1: + outlier corrector + mean, converted and written into the
// volume `vol`; mode 2: compared with the source volume `vol`, differences above tol appended to
// `sink` (unordered).
void launch_dwt_fused_inverse(const SrcVol& vol, int mode, const ChunkDev* d_chunks, const int* d_ids,
                              int nids, uint32_t nx, uint32_t ny, uint32_t nz, double tol,
                              const OutlierSink& sink, const CorrectorList& cor, cudaStream_t st)
{
  const int L = can_use_dyadic(nx, ny, nz);
  long long off[8];
  fused_scratch_elems(nx, ny, nz, off);
  FusedArgs a;
  std::memset(&a, 0, sizeof(a));
  a.chunks = d_chunks;
  a.ids = d_ids;
  a.vol = vol;
  a.k = cdf_constants();
  a.tol = tol;
  a.sink = sink;
  a.cor = cor;
  for (int l = L - 1; l >= 0; l--) {
    a.lx = int(calc_approx_detail_len(nx, l)[0]);
    a.ly = int(calc_approx_detail_len(ny, l)[0]);
    a.lz = int(calc_approx_detail_len(nz, l)[0]);
    a.apx_off = l + 1 < L ? off[l + 1] : -1;
    a.out_off = l > 0 ? off[l] : -1;
    dim3 grid;
    fused_grid(a, nids, grid);
    if (l > 0 || mode == 0)
      LAUNCH(k_inv3d<0>, grid, dim3(kFThreads), 0, st, a);
    else if (mode == 1)
      LAUNCH(k_inv3d<1>, grid, dim3(kFThreads), 0, st, a);
    else
      LAUNCH(k_inv3d<2>, grid, dim3(kFThreads), 0, st, a);
  }
}

}  // namespace sperr_b200
