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
#include <type_traits>
#ifndef SPERR_EMUL
#include <cuda.h>
#endif

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
template <bool FMA>
__device__ __forceinline__ void fwd_step(const CdfC& k, FwdState& s, double e, double o, double& e2,
                                         double& o3)
{
  const double o1n = lift_add<FMA>(s.O, k.ALPHA, __dadd_rn(s.E, e));
  const double e1n = lift_add<FMA>(s.E, k.BETA, __dadd_rn(s.O1, o1n));
  const double o2n = lift_add<FMA>(s.O1, k.GAMMA, __dadd_rn(s.E1, e1n));
  e2 = lift_scale_fwd<FMA>(k, s.E1, __dadd_rn(s.O2, o2n));
  o3 = __dmul_rn(o2n, -k.INV_EPSILON);
  s.E = e; s.O = o; s.O1 = o1n; s.E1 = e1n; s.O2 = o2n;
}

struct InvState {
  double OP, E1, O1, E2;
};
template <bool FMA>
__device__ __forceinline__ void inv_step(const CdfC& k, InvState& s, double e, double o, double& x0,
                                         double& x1)
{
  const double opn = __dmul_rn(o, -k.EPSILON);
  const double e1n = lift_scale_inv<FMA>(k, e, __dadd_rn(s.OP, opn));
  const double o1n = lift_sub<FMA>(s.OP, k.GAMMA, __dadd_rn(s.E1, e1n));
  const double e2n = lift_sub<FMA>(s.E1, k.BETA, __dadd_rn(s.O1, o1n));
  x0 = s.E2;
  x1 = lift_sub<FMA>(s.O1, k.ALPHA, __dadd_rn(s.E2, e2n));
  s.OP = opn; s.E1 = e1n; s.O1 = o1n; s.E2 = e2n;
}

// In-place lifting of one line of kFNP pairs in shared memory (element i at p[i * stride]); on
// return pairs 2 .. kFNP-3 hold the transformed values (forward: e, o interleaved as they came;
// inverse: reconstructed samples), the outer ones are scratch.
// The four lifting stages are skewed by one iteration each (stage s of iteration j works on what
// stage s-1 produced in iteration j-1), so the stages of one iteration are independent instruction
// chains: same operations on the same operands as fwd_step / inv_step, more ILP per thread.
template <bool INVERSE, bool FMA>
__device__ __forceinline__ void lift_line(const CdfC& k, double* p, int stride)
{
  if (!INVERSE) {
    double eA = 0.0, oA = 0.0, eB = 0.0;            // e[j-1], o[j-1], e[j-2]
    double p1 = 0.0, p2 = 0.0, p3 = 0.0;            // o1[j-2], o1[j-3], o1[j-4]
    double q1 = 0.0, q2 = 0.0, q3 = 0.0;            // e1[j-3], e1[j-4], e1[j-5]
    double r1 = 0.0, r2 = 0.0;                      // o2[j-5], o2[j-6]
#pragma unroll
    for (int j = 0; j < kFNP + 3; j++) {
      const double e = j < kFNP ? p[(2 * j) * stride] : 0.0;
      const double o = j < kFNP ? p[(2 * j + 1) * stride] : 0.0;
      const double n_o1 = lift_add<FMA>(oA, k.ALPHA, __dadd_rn(eA, e));
      const double n_e1 = lift_add<FMA>(eB, k.BETA, __dadd_rn(p2, p1));
      const double n_o2 = lift_add<FMA>(p3, k.GAMMA, __dadd_rn(q2, q1));
      if (j >= 7) {   // pair j - 5
        p[(2 * j - 10) * stride] = lift_scale_fwd<FMA>(k, q3, __dadd_rn(r2, r1));
        p[(2 * j - 9) * stride] = __dmul_rn(r1, -k.INV_EPSILON);
      }
      eB = eA; eA = e; oA = o;
      p3 = p2; p2 = p1; p1 = n_o1;
      q3 = q2; q2 = q1; q1 = n_e1;
      r2 = r1; r1 = n_o2;
    }
  }
  else {
    double OP1 = 0.0, OP2 = 0.0;                    // O'[k-1], O'[k-2]
    double E1a = 0.0, E1b = 0.0, E1c = 0.0;         // E1[k-1], E1[k-2], E1[k-3]
    double O1a = 0.0, O1b = 0.0, O1c = 0.0;         // O1[k-3], O1[k-4], O1[k-5]
    double E2a = 0.0, E2b = 0.0;                    // E2[k-4], E2[k-5]
#pragma unroll
    for (int j = 0; j < kFNP + 3; j++) {
      const double e = j < kFNP ? p[(2 * j) * stride] : 0.0;
      const double o = j < kFNP ? p[(2 * j + 1) * stride] : 0.0;
      const double opn = __dmul_rn(o, -k.EPSILON);
      const double e1n = lift_scale_inv<FMA>(k, e, __dadd_rn(OP1, opn));
      const double o1n = lift_sub<FMA>(OP2, k.GAMMA, __dadd_rn(E1b, E1a));
      const double e2n = lift_sub<FMA>(E1c, k.BETA, __dadd_rn(O1b, O1a));
      if (j >= 7) {   // pair j - 5
        p[(2 * j - 10) * stride] = E2b;
        p[(2 * j - 9) * stride] = lift_sub<FMA>(O1c, k.ALPHA, __dadd_rn(E2b, E2a));
      }
      OP2 = OP1; OP1 = opn;
      E1c = E1b; E1b = E1a; E1a = e1n;
      O1c = O1b; O1b = O1a; O1a = o1n;
      E2b = E2a; E2a = e2n;
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

constexpr int kNPB = 3;   // sample pairs (2 planes each) a CTA transforms per step
constexpr int kPlanes = 2 * kNPB;
constexpr size_t kFusedSmem = (size_t)kPlanes * kFI * kFP * sizeof(double);
// (A cp.async variant of the inverse transform -- every thread staging the z-phase inputs of the next
// step while the CTA lifts the current one -- was measured on B200 and dropped: 13.97 ms against
// 12.69 ms; it needed two pairs per step to keep two CTAs on an SM, and the smaller step costs more
// barriers per sample than the overlap saves.)
constexpr size_t kInvSmem = kFusedSmem;

__device__ __forceinline__ unsigned long long abs_bits(double v)
{
  return (unsigned long long)__double_as_longlong(v) & 0x7fffffffffffffffull;
}

// SRC 0: float volume, 1: double volume (both minus the chunk mean), 2: compact fp64 box in scratch
template <int SRC, bool FMA>
__global__ void __launch_bounds__(kFThreads, 2) k_fwd3d(FusedArgs a)
{
  DYN_SMEM(double, tile);   // [kPlanes][kFI][kFP]
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
  // The ChunkDev lives in global memory: anything read through `ch` after a global store has to be
  // fetched again (the store might alias it), which puts a dependent load in front of every store
  // of the z phase. Local copies of the few fields used inside the loop avoid that.
  double* const coef = ch.coef;
  const unsigned cz0 = ch.z0;

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
  const float* vf = reinterpret_cast<const float*>(a.vol.ptr);
  const double* vd = reinterpret_cast<const double*>(a.vol.ptr);

  // the 4 (x, y) columns whose z lifting this thread runs, and where their outputs go
  const size_t cnx = ch.nx, cnxy = (size_t)ch.nx * ch.ny;
  const int x = X0 + lane;
  const int xo = (x >> 1) + ((x & 1) ? ax : 0);
  FwdState st[4];
  size_t dpos[4];      // offset of the column inside a z plane of coef (de-interleaved position)
  long long apos[4];   // approx columns (x, y even): offset inside a z plane of the approx box; else -1
  bool live[4];
  for (int c = 0; c < 4; c++) {
    st[c] = FwdState{0.0, 0.0, 0.0, 0.0, 0.0};
    const int y = Y0 + warp + 8 * c;
    live[c] = x < lx && y < ly;
    dpos[c] = (size_t)((y >> 1) + ((y & 1) ? ay : 0)) * cnx + xo;
    apos[c] = ((x | y) & 1) ? -1ll
                            : (a.apx_off >= 0 ? (long long)(y >> 1) * ax + (x >> 1) : (long long)dpos[c]);
  }
  double* const abox = a.apx_off >= 0 ? ch.scratch + a.apx_off : coef;
  ASSUME_GLOBAL(coef);
  ASSUME_GLOBAL(abox);
  if (SRC == 2)   // (the hint is only given for pointers that are dereferenced: never for a null one)
    ASSUME_GLOBAL(sbox);
  const size_t aplane = a.apx_off >= 0 ? (size_t)ax * ay : cnxy;
  unsigned long long vmax = 0;

  for (int j0 = k0 - 2; j0 <= k1 + 1; j0 += kNPB) {
    // ---- load planes 2 j0 .. 2 j0 + kPlanes - 1 (mirrored) ----
    for (int p = 0; p < kPlanes; p++) {
      const int gz = mirror(2 * j0 + p, lz);
      const unsigned long long zoff = (SRC == 2 ? (unsigned long long)gz : (unsigned long long)(cz0 + gz)) * plane;
      double* const tp = tile + (size_t)p * kFI * kFP;
#ifndef SPERR_EMUL
      if (SRC != 2 && j0 + kNPB <= k1 + 1) {
        // the planes of the next step: pull them into L2 now, their loads then see L2 latency
        const int gz2 = mirror(2 * (j0 + kNPB) + p, lz);
        const unsigned long long zoff2 = (unsigned long long)(cz0 + gz2) * plane;
#pragma unroll
        for (int s = 0; s < kPer - 1; s++) {
          const void* ptr = SRC == 0 ? (const void*)(vf + zoff2 + goff[s]) : (const void*)(vd + zoff2 + goff[s]);
          asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
        }
      }
#endif
#pragma unroll
      for (int s = 0; s < kPer; s++) {
        if (s < kPer - 1 || tid + s * kFThreads < kFI * kFI) {
          double v;
          if (SRC == 0)
            v = __dsub_rn(double(__ldg(vf + zoff + goff[s])), mean);
          else if (SRC == 1)
            v = __dsub_rn(__ldg(vd + zoff + goff[s]), mean);
          else
            v = sbox[zoff + goff[s]];
          tp[sidx[s]] = v;
        }
      }
    }
    __syncthreads();
    // ---- rows (x): kPlanes * 40 lines ----
    if (tid < kPlanes * kFI)
      lift_line<false, FMA>(k, tile + (size_t)(tid / kFI) * kFI * kFP + (tid % kFI) * kFP, 1);
    __syncthreads();
    // ---- columns (y), only the x positions that are valid after the row pass ----
    if (tid < kPlanes * kFT)
      lift_line<false, FMA>(k, tile + (size_t)(tid / kFT) * kFI * kFP + kFH + tid % kFT, kFP);
    __syncthreads();
    // ---- z: streaming state in registers, 4 (x, y) columns per thread, kNPB pairs in order ----
#pragma unroll
    for (int q = 0; q < kNPB; q++) {
      const int kk = j0 + q - 2;   // output pair index
      const bool emit = kk >= k0 && kk < k1;
      const double* const te = tile + (size_t)(2 * q) * kFI * kFP + kFH * kFP + kFH + lane;
      const double* const to = te + kFI * kFP;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int ry = warp + 8 * c;
        double e2, o3;
        fwd_step<FMA>(k, st[c], te[ry * kFP], to[ry * kFP], e2, o3);
        if (emit && live[c]) {
          if (apos[c] >= 0) {   // approx band of this level (plane kk of the low band)
            abox[(size_t)kk * aplane + apos[c]] = e2;
            if (a.last) {
              const unsigned long long b = abs_bits(e2);
              vmax = b > vmax ? b : vmax;
            }
          }
          else {
            coef[(size_t)kk * cnxy + dpos[c]] = e2;
            const unsigned long long b = abs_bits(e2);
            vmax = b > vmax ? b : vmax;
          }
          if (kk < lz / 2) {   // odd z output (plane az + kk)
            coef[(size_t)(az + kk) * cnxy + dpos[c]] = o3;
            const unsigned long long b = abs_bits(o3);
            vmax = b > vmax ? b : vmax;
          }
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

#ifndef SPERR_EMUL
// ---------------------------------------------------------------------------------------------
// Level 0 of the forward transform with the source planes staged by TMA.
//
// The tile kernel above is bound by the latency of its own phases: load -> barrier -> rows ->
// barrier -> columns -> barrier -> z, with the loads of a step issued only when the previous step is
// done. Here the (40 x 40) float boxes of a step are fetched by the TMA unit
// (cp.async.bulk.tensor, completion on an mbarrier) into one of two staging buffers while the CTA
// lifts the step before: the tensor map describes the whole source volume, so a box that leaves it
// is zero-filled by the hardware and a box that leaves the CHUNK brings values of the neighbour;
// both are replaced when the staged floats are converted into the fp64 tile, because whole-sample
// mirroring x[-k] = x[k] only ever needs samples that sit in the same box (z: the mirrored plane is
// fetched instead). Two sample pairs (4 planes) per step keep two CTAs on an SM.
// ---------------------------------------------------------------------------------------------
constexpr int kTNPB = 2;
constexpr int kTPlanes = 2 * kTNPB;
constexpr int kTStage = kFI * kFI;                                   // floats of one staged plane
constexpr size_t kTmaTileBytes = (size_t)kTPlanes * kFI * kFP * sizeof(double);
constexpr size_t kTmaStageBytes = (size_t)kTPlanes * kTStage * sizeof(float);
constexpr size_t kTmaSmem = kTmaTileBytes + 2 * kTmaStageBytes + 128;   // + alignment slack

__device__ __forceinline__ unsigned smem_u32(const void* p)
{
  return unsigned(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
  // bounded: a transfer that never completes (a bad tensor map) must fail the launch, not hang it
  for (unsigned spins = 0;; spins++) {
    unsigned done;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done)
      return;
    if (spins > (1u << 24))
      __trap();
  }
}
// one (kFI x kFI x 1) box of the float volume: x fastest
__device__ __forceinline__ void tma_load_plane(void* dst, const CUtensorMap* tm, int x, int y, int z,
                                               unsigned long long* bar)
{
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
      ::"r"(smem_u32(dst)), "l"(tm), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
      : "memory");
}

template <bool FMA>
__global__ void __launch_bounds__(kFThreads, 2) k_fwd3d_tma(FusedArgs a, const __grid_constant__ CUtensorMap tmap)
{
  extern __shared__ unsigned char tma_raw_smem[];
  // 128-byte aligned start, as an OFFSET into the shared array: a pointer rebuilt from an integer
  // (round 2's first version) made every access to the staging buffers and the tile a generic LD / ST
  unsigned char* const sbase = tma_raw_smem + ((128u - (smem_u32(tma_raw_smem) & 127u)) & 127u);
  float* const stage = reinterpret_cast<float*>(sbase);                             // [2][kTPlanes][kFI * kFI]
  double* const tile = reinterpret_cast<double*>(sbase + 2 * kTmaStageBytes);       // [kTPlanes][kFI][kFP]
  __shared__ unsigned long long s_max;
  __shared__ __align__(8) unsigned long long s_bar[2];
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
  const CdfC k = a.k;
  const double mean = ch.mean;
  double* const coef = ch.coef;
  const int gx0 = int(ch.x0) + X0 - kFH, gy0 = int(ch.y0) + Y0 - kFH, gz0 = int(ch.z0);
  if (tid == 0) {
    s_max = 0;
    mbar_init(&s_bar[0], 1);
    mbar_init(&s_bar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  // step t covers pairs j0 = k0 - 2 + kTNPB t; its planes go to staging buffer t & 1
  const int nsteps = (k1 + 1 - (k0 - 2)) / kTNPB + 1;
  auto issue = [&](int t) {   // thread 0
    const int j0 = k0 - 2 + kTNPB * t;
    unsigned long long* const bar = &s_bar[t & 1];
    mbar_expect_tx(bar, unsigned(kTmaStageBytes));
    float* const dst = stage + (size_t)(t & 1) * kTPlanes * kTStage;
#pragma unroll
    for (int p = 0; p < kTPlanes; p++)
      tma_load_plane(dst + p * kTStage, &tmap, gx0, gy0, gz0 + mirror(2 * j0 + p, lz), bar);
  };
  if (tid == 0) {
    issue(0);
    if (nsteps > 1)
      issue(1);
  }

  // what this thread converts of every plane: tile elements tid, tid + 256, ...; the source of an
  // element outside the chunk is its mirror image inside the same box
  constexpr int kPer = (kFI * kFI + kFThreads - 1) / kFThreads;   // 7
  unsigned short sidx[kPer], ssrc[kPer];
  for (int s = 0; s < kPer; s++) {
    const int idx = min(tid + s * kFThreads, kFI * kFI - 1);
    const int ty = idx / kFI, tx = idx % kFI;
    sidx[s] = (unsigned short)(ty * kFP + tx);
    const int mx = mirror(X0 - kFH + tx, lx) - (X0 - kFH), my = mirror(Y0 - kFH + ty, ly) - (Y0 - kFH);
    ssrc[s] = (unsigned short)(min(max(my, 0), kFI - 1) * kFI + min(max(mx, 0), kFI - 1));
  }

  // the 4 (x, y) columns whose z lifting this thread runs, and where their outputs go
  const size_t cnx = ch.nx, cnxy = (size_t)ch.nx * ch.ny;
  const int x = X0 + lane;
  const int xo = (x >> 1) + ((x & 1) ? ax : 0);
  FwdState st[4];
  unsigned dpos[4];    // offset of the column inside a z plane of coef (a chunk holds < 2^31 values)
  unsigned apos[4];    // approx columns (x, y even): offset inside a z plane of the approx box
  unsigned live = 0, isapx = 0;   // bit c: column c is inside the box / is an approx column
  for (int c = 0; c < 4; c++) {
    st[c] = FwdState{0.0, 0.0, 0.0, 0.0, 0.0};
    const int y = Y0 + warp + 8 * c;
    if (x < lx && y < ly)
      live |= 1u << c;
    dpos[c] = unsigned((size_t)((y >> 1) + ((y & 1) ? ay : 0)) * cnx + xo);
    apos[c] = dpos[c];
    if (!((x | y) & 1)) {
      isapx |= 1u << c;
      if (a.apx_off >= 0)
        apos[c] = unsigned((y >> 1) * ax + (x >> 1));
    }
  }
  double* const abox = a.apx_off >= 0 ? ch.scratch + a.apx_off : coef;
  ASSUME_GLOBAL(coef);
  ASSUME_GLOBAL(abox);
  const unsigned cnxy32 = unsigned(cnxy), aplane32 = a.apx_off >= 0 ? unsigned(ax * ay) : cnxy32;
  const bool last_level = a.last != 0;
  const int nhigh = lz / 2;   // pairs that have a high-band sample
  unsigned long long vmax = 0;

  for (int t = 0; t < nsteps; t++) {
    const int j0 = k0 - 2 + kTNPB * t;
    // ---- the step's planes have landed: convert (float - mean, mirrored) into the fp64 tile ----
    mbar_wait(&s_bar[t & 1], unsigned(t >> 1) & 1u);
    const float* const sp = stage + (size_t)(t & 1) * kTPlanes * kTStage;
#pragma unroll
    for (int p = 0; p < kTPlanes; p++) {
      double* const tp = tile + (size_t)p * kFI * kFP;
#pragma unroll
      for (int s = 0; s < kPer; s++)
        if (s < kPer - 1 || tid + s * kFThreads < kFI * kFI)
          tp[sidx[s]] = __dsub_rn(double(sp[p * kTStage + ssrc[s]]), mean);
    }
    // every thread's reads of the staging buffer are ordered before the TMA writes that refill it
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (tid == 0 && t + 2 < nsteps)
      issue(t + 2);
    // ---- rows (x) ----
    if (tid < kTPlanes * kFI)
      lift_line<false, FMA>(k, tile + (size_t)(tid / kFI) * kFI * kFP + (tid % kFI) * kFP, 1);
    __syncthreads();
    // ---- columns (y), only the x positions that are valid after the row pass ----
    if (tid < kTPlanes * kFT)
      lift_line<false, FMA>(k, tile + (size_t)(tid / kFT) * kFI * kFP + kFH + tid % kFT, kFP);
    __syncthreads();
    // ---- z: streaming state in registers, 4 (x, y) columns per thread ----
#pragma unroll
    for (int q = 0; q < kTNPB; q++) {
      const int kk = j0 + q - 2;   // output pair index
      const bool emit = kk >= k0 && kk < k1;
      const double* const te = tile + (size_t)(2 * q) * kFI * kFP + kFH * kFP + kFH + lane;
      const double* const to = te + kFI * kFP;
#pragma unroll
      for (int c = 0; c < 4; c++) {
        const int ry = warp + 8 * c;
        double e2, o3;
        fwd_step<FMA>(k, st[c], te[ry * kFP], to[ry * kFP], e2, o3);
        if (emit && ((live >> c) & 1u)) {
          // (the approximation band of a level that is not the last one is transformed again: its
          // values do not count towards the largest coefficient)
          const bool apx = (isapx >> c) & 1u;
          if (apx)
            abox[unsigned(kk) * aplane32 + apos[c]] = e2;
          else
            coef[unsigned(kk) * cnxy32 + dpos[c]] = e2;
          if (!apx || last_level) {
            const unsigned long long b = abs_bits(e2);
            vmax = b > vmax ? b : vmax;
          }
          if (kk < nhigh) {
            coef[unsigned(az + kk) * cnxy32 + dpos[c]] = o3;
            const unsigned long long b = abs_bits(o3);
            vmax = b > vmax ? b : vmax;
          }
        }
      }
    }
    __syncthreads();
  }
  for (int o = 16; o; o >>= 1) {
    const unsigned long long t2 = __shfl_xor_sync(0xffffffffu, vmax, o);
    vmax = t2 > vmax ? t2 : vmax;
  }
  if (lane == 0 && vmax)
    atomicMax(&s_max, vmax);
  __syncthreads();
  if (tid == 0 && s_max)
    atomicMax(const_cast<unsigned long long*>(&ch.max_bits), s_max);
}
#endif   // !SPERR_EMUL

// OUT 0: fp64 box in scratch (levels > 0) or, at level 0, raw fp64 values into the volume `vol`,
//     1: destination volume (+ outlier corrector, + mean, float or double),
//     2: compare with the source volume and record the outliers
//
// Thread roles (kIThreads = 416 = 13 warps; round 2: the first version spent its time issuing
// instructions, ~3900 per thread and step with the z states spilled to local memory -- see DESIGN.md):
//   z phase   thread t < 400 owns the 2 x 2 quad (rows 2i, 2i+1; columns 2j, 2j+1), i = t / 20,
//             j = t % 20, of the 40 x 40 tile: one column of every (x, y) parity class, so which
//             sub-band a load comes from is known at compile time and the four lifting states
//             (16 doubles) stay in registers;
//   y, x      one thread per line, as in the forward kernel;
//   epilogue  warp w < 12 owns 16 rows of ONE plane of the step (plane w / 2, rows 16 (w % 2) ..),
//             lane = x: addresses are a base per step plus compile-time multiples of the row pitch.
// (out of line: a binary search that a handful of values per chunk take; inlined sixteen times it
// was half of the kernel's code)
__device__ __noinline__ double corrector_value(const CorrectorList l, unsigned chunk, unsigned pos)
{
  return corrector_lookup(l, chunk, pos);
}

// SPERR_INV_PREFETCH=1: L2 prefetch of the next step's z-phase inputs. Measured on B200 (1024^3):
// d.idwt 9.34 ms with it, 9.23 ms without -- two CTAs per SM already overlap one CTA's loads with
// the other's lifting -- so it is off.
#ifndef SPERR_INV_PREFETCH
#define SPERR_INV_PREFETCH 0
#endif
constexpr int kIThreads = 416;
constexpr int kIQuads = kFNP * kFNP;   // 400
static_assert(kIQuads <= kIThreads && kPlanes * kFI <= kIThreads && 2 * kPlanes <= kIThreads / 32, "roles");
constexpr int kIRows = 16;             // rows of a plane an epilogue warp owns

template <int OUT, bool FMA, bool F32>
__global__ void __launch_bounds__(kIThreads, 2) k_inv3d(FusedArgs a)
{
  DYN_SMEM(double, tile);   // [kPlanes][kFI][kFP]
  const unsigned cid = a.ids[blockIdx.y];
  const ChunkDev& ch = a.chunks[cid];
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
  // local copies of what the loop reads through `ch` (see k_fwd3d)
  const double* const coef = ch.coef;
  const unsigned cz0 = ch.z0;
  const double* const abox = a.apx_off >= 0 ? ch.scratch + a.apx_off : coef;
  ASSUME_GLOBAL(coef);
  ASSUME_GLOBAL(abox);
  const unsigned cnxy32 = unsigned(cnxy), aplane32 = a.apx_off >= 0 ? unsigned(ax * ay) : cnxy32;

  // ---- z phase: my quad. Slot s = 2 py + px is the column (2i + py, 2j + px) of the tile; the tile
  // starts at an even sample and mirroring keeps parity, so slot s is always of parity class
  // (px, py): slot 0 reads its even-z samples from the approximation box, the others from coef.
  const bool zthread = tid < kIQuads;
  const int qi = zthread ? tid / kFNP : 0, qj = zthread ? tid % kFNP : 0;
  unsigned coff[4];       // offset inside a z plane of coef (chunks hold < 2^31 values)
  unsigned eoff0 = 0;     // slot 0: offset inside a z plane of the approximation box
  InvState st[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    st[s] = InvState{0.0, 0.0, 0.0, 0.0};
    const int gx = mirror(X0 - kFH + 2 * qj + (s & 1), lx), gy = mirror(Y0 - kFH + 2 * qi + (s >> 1), ly);
    const int xo = (gx >> 1) + ((gx & 1) ? ax : 0), yo = (gy >> 1) + ((gy & 1) ? ay : 0);
    coff[s] = unsigned((size_t)yo * cnx + xo);
    if (s == 0)   // (a position clamped by mirror() can have the wrong parity: its value is never used,
                  // the load only has to stay inside the box)
      eoff0 = ((gx | gy) & 1) ? 0u
                              : (a.apx_off >= 0 ? unsigned((gy >> 1) * ax + (gx >> 1))
                                                : unsigned((size_t)(gy >> 1) * cnx + (gx >> 1)));
  }
  double* const zt = tile + (2 * qi) * kFP + 2 * qj;   // my quad in plane 0 of the tile
#if SPERR_INV_PREFETCH
  const bool pf_thread = zthread && ((qj & 3) == 0 || qj == kFNP - 1);
#endif

  // ---- epilogue: my plane of the step and my rows in it ----
  const int lane = tid & 31, warp = tid >> 5;
  const int ep_p = warp >> 1, ep_r0 = (warp & 1) * kIRows;
  const bool ep_warp = warp < 2 * kPlanes;
  const bool ep_x = X0 + lane < lx;
  const int ep_rows = ep_warp ? max(0, min(kIRows, ly - Y0 - ep_r0)) : 0;   // rows of mine inside the box
  const double* const ep_t = tile + (size_t)ep_p * kFI * kFP + (kFH + ep_r0) * kFP + kFH + lane;
  const unsigned long long vplane = a.vol.vx * a.vol.vy;
  const double ep_mean = ch.mean;
  // offset of (lane, first row of mine) inside a z plane of the volume / of the chunk
  const unsigned long long ep_g0 = (unsigned long long)(ch.y0 + Y0 + ep_r0) * a.vol.vx + (ch.x0 + X0 + lane);
  const unsigned ep_c0 = unsigned((size_t)(Y0 + ep_r0) * cnx + X0 + lane);
  double* const obox = (OUT == 0 && a.out_off >= 0) ? ch.scratch + a.out_off : nullptr;
  // mode 1 with correctors: lane r < 16 fetches the 32 corrector flags of row r of the warp's plane
  // at the top of the step (in flight during the lifting passes); the epilogue gets them by shuffle
  const uint32_t* const obits = (OUT == 1 && a.cor.key) ? ch.obits : nullptr;
  const unsigned fl_i0 = unsigned((size_t)(Y0 + ep_r0 + (lane & (kIRows - 1))) * cnx + X0);

  for (int j0 = k0 - 2; j0 <= k1 + 1; j0 += kNPB) {
    // plane of mine in this step: pair kk = j0 + p / 2 - 2, sample z = 2 kk + p % 2
    const int ep_kk = j0 + (ep_p >> 1) - 2;
    const int ep_z = 2 * ep_kk + (ep_p & 1);
    const bool ep_on = ep_warp && ep_kk >= k0 && ep_kk < k1 && ep_z < lz;   // warp-uniform
    unsigned flags = 0;
    if (OUT == 1 && obits && ep_on && lane < ep_rows) {
      const unsigned i0 = unsigned(ep_z) * cnxy32 + fl_i0;
      const unsigned sh = i0 & 31u;
      unsigned long long w = __ldg(obits + (i0 >> 5));
      if (sh)   // the row straddles two words (every chunk's flag array ends with a spare word)
        w |= (unsigned long long)__ldg(obits + (i0 >> 5) + 1) << 32;
      flags = unsigned(w >> sh);
    }
#ifndef SPERR_EMUL
    if (OUT == 2 && ep_on && lane < ep_rows) {
      // the source values the epilogue compares with: lane r pulls row r's line into L2
      const unsigned long long e = (unsigned long long)(cz0 + ep_z) * vplane + ep_g0 - lane + (unsigned long long)lane * a.vol.vx;
      const char* const ptr = reinterpret_cast<const char*>(a.vol.ptr) + e * (F32 ? 4 : 8);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
      if (!F32)
        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr + 128));
    }
#endif
    // ---- z: pair j = (low-band plane mirror(2j) / 2, high-band plane az + mirror(2j + 1) / 2) ----
#if SPERR_INV_PREFETCH && !defined(SPERR_EMUL)
    // the inputs of the NEXT step: every fourth quad of a quad row pulls the lines of its eight loads
    // into L2 now (a quad row reads 20 consecutive doubles per sub-band: the lines of quads 0, 4, ..
    // 16 and of the last one cover them), so that the loads of the next step see L2 latency
    if (pf_thread && j0 + kNPB <= k1 + 1) {
#pragma unroll
      for (int q = 0; q < kNPB; q++) {
        const int j = j0 + kNPB + q;
        const unsigned ze = unsigned(mirror(2 * j, lz) >> 1), zo = unsigned(az + (mirror(2 * j + 1, lz) >> 1));
        const unsigned ie = ze * cnxy32, io = zo * cnxy32, ia = ze * aplane32;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(abox + (ia + eoff0)));
#pragma unroll
        for (int s = 1; s < 4; s++)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(coef + (ie + coff[s])));
#pragma unroll
        for (int s = 0; s < 4; s++)
          asm volatile("prefetch.global.L2 [%0];" ::"l"(coef + (io + coff[s])));
      }
    }
#endif
    if (zthread) {
#pragma unroll
      for (int q = 0; q < kNPB; q++) {
        const int j = j0 + q;
        // plane offsets as 32-bit element indices (a chunk holds < 2^31 values): one add and one
        // widening multiply-add per load instead of 64-bit arithmetic per thread
        const unsigned ze = unsigned(mirror(2 * j, lz) >> 1), zo = unsigned(az + (mirror(2 * j + 1, lz) >> 1));
        const unsigned ie = ze * cnxy32, io = zo * cnxy32, ia = ze * aplane32;
        double ev[4], ov[4];
        ev[0] = abox[ia + eoff0];
#pragma unroll
        for (int s = 1; s < 4; s++)
          ev[s] = coef[ie + coff[s]];
#pragma unroll
        for (int s = 0; s < 4; s++)
          ov[s] = coef[io + coff[s]];
        double* const t0 = zt + (size_t)(2 * q) * kFI * kFP;
#pragma unroll
        for (int s = 0; s < 4; s++) {
          double x0, x1;
          inv_step<FMA>(k, st[s], ev[s], ov[s], x0, x1);
          t0[(s >> 1) * kFP + (s & 1)] = x0;
          t0[kFI * kFP + (s >> 1) * kFP + (s & 1)] = x1;
        }
      }
    }
    __syncthreads();
    // planes 2 (j0 - 2) .. of the rebuilt box sit in the tile; pair q is wanted iff k0 <= j0+q-2 < k1
    // ---- columns (y) over all x of the tile, then rows (x) over the valid y ----
    if (tid < kPlanes * kFI)
      lift_line<true, FMA>(k, tile + (size_t)(tid / kFI) * kFI * kFP + tid % kFI, kFP);
    __syncthreads();
    if (tid < kPlanes * kFT)
      lift_line<true, FMA>(k, tile + (size_t)(tid / kFT) * kFI * kFP + (kFH + tid % kFT) * kFP, 1);
    __syncthreads();
    // ---- epilogue: 16 rows x 32 values of one plane per warp. The common case (my rows and my x
    // inside the box, no corrector in my rows) is a straight run of load / convert / store with
    // compile-time offsets; everything else takes the general loops. ----
    if (ep_on) {
      const bool whole = ep_rows == kIRows && X0 + kFT <= lx;   // warp-uniform
      if (OUT == 0 && obox) {
        double* const ob = obox + ((size_t)ep_z * ly + Y0 + ep_r0) * lx + X0 + lane;
        ASSUME_GLOBAL(ob);
        if (whole) {
#pragma unroll
          for (int r = 0; r < kIRows; r++)
            ob[r * lx] = ep_t[r * kFP];
        }
        else if (ep_x) {
          for (int r = 0; r < ep_rows; r++)
            ob[r * lx] = ep_t[r * kFP];
        }
      }
      else if (OUT == 0) {
        double* const vd = reinterpret_cast<double*>(const_cast<void*>(a.vol.ptr)) + (unsigned long long)(cz0 + ep_z) * vplane + ep_g0;
        if (ep_x)
          for (int r = 0; r < ep_rows; r++)
            vd[(unsigned long long)r * a.vol.vx] = ep_t[r * kFP];
      }
      else if (OUT == 1) {
        typedef typename std::conditional<F32, float, double>::type TOut;
        TOut* const vo = reinterpret_cast<TOut*>(const_cast<void*>(a.vol.ptr)) + (unsigned long long)(cz0 + ep_z) * vplane + ep_g0;
        const unsigned long long vx = a.vol.vx;
        const bool marked = obits && __any_sync(0xffffffffu, flags != 0);
        if (whole && !marked) {
#pragma unroll
          for (int r = 0; r < kIRows; r++) {
            const double w = __dadd_rn(ep_t[r * kFP], ep_mean);
            vo[r * vx] = F32 ? TOut(__double2float_rn(w)) : TOut(w);
          }
        }
        else {
          const unsigned iz = unsigned(ep_z) * cnxy32 + ep_c0;
          for (int r = 0; r < kIRows; r++) {   // (all lanes take part in the shuffle)
            const unsigned rowflags = __shfl_sync(0xffffffffu, flags, r);
            if (r < ep_rows && ep_x) {
              double w = ep_t[r * kFP];
              if ((rowflags >> lane) & 1u)   // src/SPECK_FLT.cpp:576-585: correctors are added before the mean
                w = __dadd_rn(w, corrector_value(a.cor, cid, iz + unsigned(r) * unsigned(cnx)));
              w = __dadd_rn(w, ep_mean);
              vo[r * vx] = F32 ? TOut(__double2float_rn(w)) : TOut(w);
            }
          }
        }
      }
      else {
        typedef typename std::conditional<F32, float, double>::type TIn;
        const TIn* const vi = reinterpret_cast<const TIn*>(a.vol.ptr) + (unsigned long long)(cz0 + ep_z) * vplane + ep_g0;
        const unsigned long long vx = a.vol.vx;
        const unsigned iz = unsigned(ep_z) * cnxy32 + ep_c0;
        if (whole) {
          // the source values first (independent loads in flight), then the comparisons
#pragma unroll
          for (int h = 0; h < kIRows; h += 8) {
            double orig[8];
#pragma unroll
            for (int r = 0; r < 8; r++)
              orig[r] = double(__ldg(vi + (h + r) * vx));
#pragma unroll
            for (int r = 0; r < 8; r++) {
              const double diff = __dsub_rn(__dsub_rn(orig[r], ep_mean), ep_t[(h + r) * kFP]);
              if (fabs(diff) > a.tol)
                outlier_append(a.sink, cid, iz + unsigned(h + r) * unsigned(cnx), diff);
            }
          }
        }
        else if (ep_x) {
          for (int r = 0; r < ep_rows; r++) {
            const double diff = __dsub_rn(__dsub_rn(double(__ldg(vi + r * vx)), ep_mean), ep_t[r * kFP]);
            if (fabs(diff) > a.tol)
              outlier_append(a.sink, cid, iz + unsigned(r) * unsigned(cnx), diff);
          }
        }
      }
    }
    __syncthreads();
  }
}

// Multi-resolution output (CDF97::idwt3d_multi_res, /root/reference/src/CDF97.cpp:150-168;
// SPECK_FLT::decompress :592-603; SPERR3D_OMP_D.cpp:70-126): the approximation box a chunk holds
// before transform level `lev` is undone, plus the chunk mean, placed at the chunk's position in
// the coarsened volume. After the fused inverse transform those boxes are still there: level L in
// the corner of `coef`, levels 1 .. L-1 compact in `scratch`.
__global__ void k_level_gather(const ChunkDev* chunks, int ax, int ay, int az, long long src_off,
                               void* dst, int is_float, unsigned long long vx, unsigned long long vy)
{
  const ChunkDev& ch = chunks[blockIdx.y];
  const unsigned long long n = (unsigned long long)ax * ay * az;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  // position of the chunk in the chunk grid (the volume is divisible by the chunk extents)
  const unsigned long long ox = (unsigned long long)(ch.x0 / ch.nx) * ax, oy = (unsigned long long)(ch.y0 / ch.ny) * ay,
                           oz = (unsigned long long)(ch.z0 / ch.nz) * az;
  for (unsigned long long e = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += stride) {
    const unsigned x = unsigned(e % ax), y = unsigned((e / ax) % ay), z = unsigned(e / ((unsigned long long)ax * ay));
    double v;
    if (ch.is_const)
      v = ch.first_val;   // the reference leaves such a chunk's coarse boxes unwritten
    else {
      const double c = src_off >= 0 ? ch.scratch[src_off + e]
                                    : ch.coef[((size_t)z * ch.ny + y) * ch.nx + x];
      v = __dadd_rn(c, ch.mean);
    }
    const unsigned long long g = ((oz + z) * vy + (oy + y)) * vx + (ox + x);
    if (is_float)
      reinterpret_cast<float*>(dst)[g] = __double2float_rn(v);
    else
      reinterpret_cast<double*>(dst)[g] = v;
  }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------

// Coarse level `h` (0 = coarsest) of chunks of shape nx x ny x nz: see k_level_gather. `dst` is the
// coarsened volume, vx x vy values per plane.
void launch_level_gather(const ChunkDev* d_chunks, int nchunks, uint32_t nx, uint32_t ny, uint32_t nz,
                         int h, void* dst, int is_float, size_t vx, size_t vy, cudaStream_t st)
{
  const int L = can_use_dyadic(nx, ny, nz);
  const int lev = L - h;
  long long off[8];
  fused_scratch_elems(nx, ny, nz, off);
  const int ax = int(calc_approx_detail_len(nx, lev)[0]), ay = int(calc_approx_detail_len(ny, lev)[0]),
            az = int(calc_approx_detail_len(nz, lev)[0]);
  const size_t n = size_t(ax) * ay * az;
  const unsigned gx = unsigned(std::min<size_t>((n + 255) / 256, 1024));
  LAUNCH(k_level_gather, dim3(gx, unsigned(nchunks)), dim3(256), 0, st, d_chunks, ax, ay, az,
         lev == L ? -1ll : off[lev], dst, is_float, (unsigned long long)vx, (unsigned long long)vy);
}

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

#ifndef SPERR_EMUL
// Tensor map of the whole float source volume (x fastest), boxes of kFI x kFI x 1 samples.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_volume_tmap(const SrcVol& src, CUtensorMap& tm)
{
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
    else
      cudaGetLastError();
  }
  if (!fn)
    return false;
  const cuuint64_t gdim[3] = {src.vx, src.vy, src.vz};
  const cuuint64_t gstr[2] = {src.vx * 4ull, src.vx * src.vy * 4ull};
  const cuuint32_t box[3] = {kFI, kFI, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  return fn(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(src.ptr), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
// TMA needs a 16-byte aligned base and row pitch; mirroring inside the box needs the box to reach
// 4 samples past an edge it straddles (any level-0 box of at least 9 samples does).
static bool fwd_tma_usable(const SrcVol& src, const FusedArgs& a)
{
  if (std::getenv("SPERR_B200_NO_TMA"))
    return false;
  // (row pitch a multiple of 128 bytes: with a 400-byte pitch -- a 100 x 88 x 40 volume -- the copy
  // instruction itself faulted on B200, "illegal instruction" at the UTMALDG, although the tensor
  // map had been accepted; such volumes take the tile kernel)
  return src.is_float && src.vz > 0 && (reinterpret_cast<uintptr_t>(src.ptr) & 127) == 0 && src.vx % 32 == 0 &&
         a.lx >= 2 * kFH + 1 && a.ly >= 2 * kFH + 1 && a.lz >= 2 * kFH + 1;
}
#endif

static void fused_attrs()
{
#ifndef SPERR_EMUL
  static rt::OncePerDevice once;
  if (!once.first())
    return;
  const int sm = int(kFusedSmem);
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  const int smi = int(kInvSmem);
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<0, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi));
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<0, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi));
#define SPERR_INV_ATTR(O) \
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<O, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi)); \
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<O, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi));  \
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<O, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi));  \
  RT_CHECK(cudaFuncSetAttribute(k_inv3d<O, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smi));
  SPERR_INV_ATTR(1)
  SPERR_INV_ATTR(2)
#undef SPERR_INV_ATTR
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d<2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTmaSmem)));
  RT_CHECK(cudaFuncSetAttribute(k_fwd3d_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(kTmaSmem)));
#endif
}

// Forward transform of dyadic chunks straight from the source volume: coef receives the final
// coefficients, ChunkDev::max_bits the largest magnitude.
void launch_dwt_fused_forward(const SrcVol& src, const ChunkDev* d_chunks, const int* d_ids, int nids,
                              uint32_t nx, uint32_t ny, uint32_t nz, cudaStream_t st)
{
  fused_attrs();
  const int L = can_use_dyadic(nx, ny, nz);
  long long off[8];
  fused_scratch_elems(nx, ny, nz, off);
  FusedArgs a{};
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
    const int which = l > 0 ? 2 : (src.is_float ? 0 : 1);
#ifndef SPERR_EMUL
    if (which == 0 && fwd_tma_usable(src, a)) {
      CUtensorMap tm;
      if (make_volume_tmap(src, tm)) {
        if (a.k.fma)
          LAUNCH((k_fwd3d_tma<true>), grid, dim3(kFThreads), kTmaSmem, st, a, tm);
        else
          LAUNCH((k_fwd3d_tma<false>), grid, dim3(kFThreads), kTmaSmem, st, a, tm);
        continue;
      }
    }
#endif
#define SPERR_FWD(S, F) LAUNCH((k_fwd3d<S, F>), grid, dim3(kFThreads), kFusedSmem, st, a)
    if (a.k.fma) {
      if (which == 2) SPERR_FWD(2, true);
      else if (which == 0) SPERR_FWD(0, true);
      else SPERR_FWD(1, true);
    }
    else {
      if (which == 2) SPERR_FWD(2, false);
      else if (which == 0) SPERR_FWD(0, false);
      else SPERR_FWD(1, false);
    }
#undef SPERR_FWD
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
  fused_attrs();
  const int L = can_use_dyadic(nx, ny, nz);
  long long off[8];
  fused_scratch_elems(nx, ny, nz, off);
  FusedArgs a{};
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
    const int which = (l > 0 || mode == 0) ? 0 : (mode == 1 ? 1 : 2);
    if (which == 2) {
      // The outlier scan runs beside the SPECK3D encoder, whose kernels have priority (pipeline.cu):
      // short-lived CTAs (about 32 sample pairs each instead of the whole z extent, 4 of them
      // recomputed) hand their slots over quickly.
      static const int seg_pairs = std::getenv("SPERR_B200_SCAN_SEG_PAIRS") ? std::atoi(std::getenv("SPERR_B200_SCAN_SEG_PAIRS")) : 0;   // off: measured, no gain
      const int az = a.lz - a.lz / 2;
      if (seg_pairs > 0) {
        int zs = a.zsegs;
        while (az / (zs * 2) >= seg_pairs)
          zs *= 2;
        a.zsegs = zs;
        grid = dim3(unsigned(a.tiles_x * a.tiles_y * zs), unsigned(nids));
      }
    }
#define SPERR_INV(O, F, T) LAUNCH((k_inv3d<O, F, T>), grid, dim3(kIThreads), kInvSmem, st, a)
#define SPERR_INV_T(O, F)        \
  do {                           \
    if (vol.is_float)            \
      SPERR_INV(O, F, true);     \
    else                         \
      SPERR_INV(O, F, false);    \
  } while (0)
    if (a.k.fma) {
      if (which == 0) SPERR_INV(0, true, false);
      else if (which == 1) SPERR_INV_T(1, true);
      else SPERR_INV_T(2, true);
    }
    else {
      if (which == 0) SPERR_INV(0, false, false);
      else if (which == 1) SPERR_INV_T(1, false);
      else SPERR_INV_T(2, false);
    }
#undef SPERR_INV_T
#undef SPERR_INV
  }
}

}  // namespace sperr_b200
