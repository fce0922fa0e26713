// Integer SPECK decoders: job description and host entry points (kernels in speck_dec.cuh).
#pragma once

#include "kernels.h"

namespace sperr_b200 {

struct DChild {
  node_t id;                 // sets
  unsigned long long idx;    // pixels: raster index
  int lis;                   // sets: list index
  int pixel;
};

// Decoder state of one chunk (device memory).
struct DecChunk {
  const uint32_t* bits;          // payload, staged 4-byte aligned and zero padded
  unsigned long long avail;      // bits really present: min(total_bits, 8 * payload bytes)
  int planes;
  int skip;                      // nothing to decode (constant chunk, or no stream)
  unsigned long long n;          // number of coefficients
  int shape;                     // 3D: index into the shape tables
  int wide;                      // magnitudes are 64-bit
  void* mag;                     // n magnitudes, zero on entry
  uint32_t* signs;               // bit i = 1: non-negative; all ones on entry
  uint32_t* lip;                 // masks, all zero on entry
  uint32_t* lsp;
  uint32_t* newm;
  uint32_t* sigarr;              // scratch of the LIP pass (ceil(n / 32) + 2 words each)
  uint32_t* signarr;
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
  unsigned long long* scr;       // per-thread append staging of the token expanders
  unsigned long long stage_words;   // words in `bits`
  // clock64() totals of the fast decoder's phases (thread 0): 0 LIP pass, 1 window staging + body
  // tables, 2 token chains, 3 token expansion + commit, 4 tree walker, 5 refinement, 6 windows built
  unsigned long long prof[8];
};

constexpr int kFastScrPerThread = 96;   // entries of DecChunk::scr per decoder thread

// One integer stream to decode. `mag` (zeroed) and `signs` (all ones) are provided by the caller.
struct DecJob {
  unsigned long long n = 0;
  int shape = 0;
  bool skip = true;
  const unsigned char* d_payload = nullptr;   // device pointer: bytes after the 9-byte header
  unsigned long long payload_bytes = 0;
  int planes = 0;
  unsigned long long total_bits = 0;
  void* mag = nullptr;
  uint32_t* signs = nullptr;
  int wide = 0;
  int nlis = 0;
  const unsigned long long* d_lis_off = nullptr;   // device, nlis + 1 entries
  unsigned long long lis_total = 0;                // total list capacity (entries)
  int pow2 = 0, Dx = 0, Dy = 0, Dz = 0;            // fast path (see DecChunk)
  unsigned nx = 0, ny = 0;
  int nroots = 0;
  unsigned long long roots[kMaxRoots] = {0};
};

struct DecWork {
  rt::DBuf dchunks, masks, lis, lis_cnt, stage, aux, scr;
  std::vector<DecChunk> h;   // copy of the device state after the last run
};

// Decodes every job; afterwards w.h[c].lsp is the final significance mask of job c.
void speck3d_decode(DecWork& w, const std::vector<DecJob>& jobs, const ShapeDev* d_shapes,
                    cudaStream_t st);
void speck1d_decode(DecWork& w, const std::vector<DecJob>& jobs, cudaStream_t st);
void launch_outlier_apply(const DecChunk* d_jobs, const ChunkDev* d_chunks, const double* d_tols,
                          int nchunks, size_t max_n, cudaStream_t st);

}  // namespace sperr_b200
