// Integer SPECK decoders: job description and host entry points (kernels in speck_dec*.cuh).
#pragma once

#include "kernels.h"

namespace sperr_b200 {

struct DChild {
  node_t id;                 // sets
  unsigned long long idx;    // pixels: raster index
  int lis;                   // sets: list index
  int pixel;
};

constexpr int kMaxPlanes = 64;

// Decoder state of one chunk (device memory).
//
// The per-chunk kernels decode only the SORTING passes: they record for every coefficient the plane
// at which it became significant (and its sign) in `pl`, and for every plane where its refinement
// bits sit in the stream. Magnitudes are rebuilt afterwards by grid-wide kernels (k_rec_*), because
// a coefficient's refinement bit in plane n is simply bit number rank_n(i) of that plane's section,
// rank_n(i) = how many coefficients before i (raster order) were significant before plane n.
struct DecChunk {
  const uint32_t* bits;          // payload, staged 4-byte aligned and zero padded
  unsigned long long avail;      // bits really present: min(total_bits, 8 * payload bytes)
  int planes;
  int skip;                      // nothing to decode (constant chunk, or no stream)
  unsigned long long n;          // number of coefficients
  int shape;                     // 3D: index into the shape tables
  int kind;                      // 0: 3D coefficient stream, 1: 1D outlier stream, 2: 2D slice
  node_t iset;                   // 2D: the live set I (0: none)
  uint8_t* pl;                   // n bytes, 0xFF on entry: plane of significance | negative << 7
  uint32_t* lip;                 // LIP mask, all zero on entry
  uint32_t* sigarr;              // scratch of the LIP pass (ceil(n / 32) + 2 words each)
  uint32_t* signarr;
  // one bit per group of 128 words (4096 pixels) of the LIP mask: set when a pixel of the group is
  // put on the list, never cleared -- the two sweeps of a LIP pass skip the groups whose bit is clear
  uint32_t* lipsum;
  node_t* lis;                   // list storage
  const unsigned long long* lis_off;   // nlis + 1 offsets into `lis`
  unsigned* lis_cnt;             // nlis counters
  int nlis;
  unsigned err;                  // 1: a list overflowed its capacity
  // fast path for power-of-two trees (speck_dec_fast.cuh): every set is an aligned box
  int pow2;
  int Dx, Dy, Dz;                // bisection depth of every axis
  unsigned nx, ny;
  int nroots;
  unsigned long long roots[kMaxRoots];   // initial sets as (depth << 32 | linear index), list order
  unsigned long long stage_words;        // words in `bits`
  // refinement sections: plane n's bits start at ref_base[n]; the first ref_cnt[n] significant
  // coefficients (raster order) own one bit each (fewer than all of them only where a truncated
  // stream ends inside the section)
  unsigned long long ref_base[kMaxPlanes], ref_cnt[kMaxPlanes];
  // thread-block cluster of R CTAs per stream (speck_dec_fast.cuh): mailbox and append scratch
  int R;
  struct ClusterBox* box;
  node_t* scr;
  // clock64() totals of the fast decoder's phases (thread 0): 0 LIP pass, 1 window staging + body
  // tables, 2 token chains, 3 token expansion, 4 tree walker, 6 windows built
  unsigned long long prof[8];
  // debugging aid (SPERR_B200_DECPROF): per plane (position, LIP population, new significant, sets in the
  // lists) after the LIP part, the chain phases and the walker
  unsigned long long dbg[kMaxPlanes][3][2];
  unsigned dbgl[2][3][32];   // the first two planes: sets per list at the three stages
  unsigned dbgw[16][8];      // the first walker list visits: plane, depth, roots, visited, survivors, ...
  unsigned dbgw_n, dbg_pad;
};

// One integer stream to decode.
struct DecJob {
  unsigned long long n = 0;
  int shape = 0;
  int kind = 0;   // 0: 3D coefficient stream, 1: 1D outlier stream, 2: 2D slice (SPECK2D)
  bool skip = true;
  const unsigned char* d_payload = nullptr;   // device pointer: bytes after the 9-byte header
  unsigned long long payload_bytes = 0;
  int planes = 0;
  unsigned long long total_bits = 0;
  int nlis = 0;
  const unsigned long long* d_lis_off = nullptr;   // device, nlis + 1 entries
  unsigned long long lis_total = 0;                // total list capacity (entries)
  int pow2 = 0, Dx = 0, Dy = 0, Dz = 0;            // fast path (see DecChunk)
  unsigned nx = 0, ny = 0;
  int nroots = 0;
  unsigned long long roots[kMaxRoots] = {0};
  int nxf2d = 0;   // 2D fast path: transform levels = initial part_level of the set I
};

struct DecWork {
  rt::DBuf dchunks, masks, pl, lis, lis_cnt, stage, aux, counts, boxes, scr;
  std::vector<DecChunk> h;   // copy of the device state after the last run
  size_t max_n = 0;        // largest decoded job
  size_t fill_n = 0;       // largest chunk of the batch (set by the caller; zero-fill of mode 0)
  int max_planes = 0;
};

// Decodes the sorting passes of every job (see DecChunk); 3D and 1D jobs may be mixed.
void speck_decode(DecWork& w, const std::vector<DecJob>& jobs, const ShapeDev* d_shapes,
                  cudaStream_t st);

// Rebuilds the magnitudes from the decoded state of speck*_decode and
//   mode 0: writes the de-quantised coefficients q * mag * sign to chunks[c].coef
//           (SPECK_FLT::m_midtread_inv_quantize, /root/reference/src/SPECK_FLT.cpp:373-399);
//   mode 1: treats them as outlier correctors with tolerance tols[c] (Outlier_Coder::
//           m_inverse_quantize, src/Outlier_Coder.cpp:206-234): appends (chunk << 32 | position,
//           corrector) to `sink` and sets the value's bit in chunks[c].obits. The consumers add the
//           corrector before the mean (src/SPECK_FLT.cpp:576-585).
// Works on jobs [first, first + count) of the last speck_decode; job first + c belongs to chunks[c].
void speck_reconstruct(DecWork& w, const ChunkDev* d_chunks, int mode, const double* d_tols,
                       int first, int count, const OutlierSink& sink, cudaStream_t st);
void launch_apply_correctors(const ChunkDev* d_chunks, const unsigned long long* d_key,
                             const double* d_val, unsigned long long n, cudaStream_t st);

}  // namespace sperr_b200
