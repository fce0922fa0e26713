// Chunk-batch pipelines: the GPU counterpart of SPERR3D_OMP_C / SPERR3D_OMP_D's per-chunk loop
// (/root/reference/src/SPERR3D_OMP_C.cpp:62-141, src/SPERR3D_OMP_D.cpp:51-135) with
// SPECK_FLT::compress / decompress (src/SPECK_FLT.cpp:401-606) inside.
#pragma once

#include <functional>
#include <mutex>

#include "batch.h"
#include "outlier.h"
#include "speck.h"
#include "speck_dec.h"

namespace sperr_b200 {

enum Mode { kModeRate = 1, kModePSNR = 2, kModePWE = 3 };

struct StageTimes {  // milliseconds, filled when profiling is requested
  float stats = 0, transform = 0, quant = 0, outlier = 0, speck = 0, assemble = 0;
};

class Compressor {
 public:
  // Compresses `chunks` of the device-resident volume `src`. On return `d_out` holds the chunk
  // streams back to back (device memory) and `lens` their lengths. Throws on error.
  void compress(const SrcVol& src, const std::vector<Chunk>& chunks, int mode, double quality,
                bool is_2d, rt::DBuf& d_out, std::vector<size_t>& lens, cudaStream_t st);
  // Batching hooks for callers that overlap transfers with the coder: at most `max_batch` chunks
  // per batch (0: as many as fit), `before_batch(first, count)` runs before a batch is touched.
  size_t max_batch = 0;
  std::function<void(size_t, size_t)> before_batch;
  // Inside a batch: when set, statistics + forward transform run `group_chunks` chunks at a time and
  // `before_group(first, count)` is called before a group's part of the source is read (it makes
  // the stream wait for that part); without grouping it is called once for the whole batch.
  size_t group_chunks = 0;
  std::function<void(size_t, size_t)> before_group;

 private:
  size_t batch_first_ = 0;
  bool whole_call_ = false;
  void run_batch(const SrcVol& src, const std::vector<Chunk>& chunks, int mode, double quality,
                 bool is_2d, std::vector<std::vector<uint8_t>>& hdrs, cudaStream_t st);
  BatchBuffers b_;
  Speck3DEncoder enc_, enc_hp_;
  OutlierCoder out_;
  rt::DBuf stride_mean_, nstrides_, not_const_, ids_, mse_ids_, mse_q_, mse_part_, mse_out_;
  rt::DBuf asm_hdr_, asm_pieces_;   // stream assembly: headers and the piece list
  // results of the last batch
  std::vector<EncResult> spk_res_, out_res_;
#ifndef SPERR_EMUL
  cudaStream_t side_ = nullptr;   // PWE: the outlier path runs beside the SPECK3D encoder
  cudaEvent_t side_ev_ = nullptr;
  // PWE: the SPECK3D encoder (the critical path) runs on a stream of the highest priority, so that
  // CTA slots the outlier path's bandwidth kernels free go to the encoder's kernels first
  cudaStream_t hi_ = nullptr;
  cudaEvent_t hi_ev_ = nullptr;
  cudaEvent_t stagger_ev_ = nullptr;   // SPERR_B200_STAGGER: the encoder has reached its bit-plane loop
#endif
};

struct ChunkStream {   // where a chunk's stream sits inside the container
  size_t off, len;
};

// Multi-resolution decoding: device buffers of the coarsened volumes, coarsest first
// (sperr::coarsened_resolutions, /root/reference/src/sperr_helper.cpp:70-123).
struct MultiRes {
  int is_float = 0;
  std::vector<void*> d_level;                 // one coarsened volume per level
  std::vector<std::array<size_t, 3>> dims;    // its extents
};

class Decompressor {
 public:
  // Decodes `chunks` (streams at h_stream + cs[i].off; d_stream is the same container in device
  // memory) into the device-resident volume `dst`. Throws on malformed input.
  void decompress(const uint8_t* h_stream, const uint8_t* d_stream, const std::vector<Chunk>& chunks,
                  const std::vector<ChunkStream>& cs, const SrcVol& dst, cudaStream_t st,
                  bool is_2d = false);
  // at most `max_batch` chunks per batch (0: as many as fit); `after_batch(first, count)` runs when
  // the values of a batch are complete in `dst`
  size_t max_batch = 0;
  std::function<void(size_t, size_t)> after_batch;
  // Inside a batch: when set, the final stage (inverse transform into `dst`) runs `group_chunks`
  // chunks at a time and `after_group(first, count, event)` is called after each group has been
  // launched; the values of those chunks are in `dst` once the event (owned by the callee) has
  // completed. `groups_posted` tells whether the last batch was reported that way.
  size_t group_chunks = 0;
#ifndef SPERR_EMUL
  std::function<void(size_t, size_t, cudaEvent_t)> after_group;
#else
  std::function<void(size_t, size_t, void*)> after_group;
#endif
  bool groups_posted = false;
  // when set, every batch also writes the coarse levels of its chunks (all chunks must have the
  // same dyadic shape and the volume must be divisible by it: the caller checks)
  const MultiRes* multires = nullptr;

 private:
  void run_batch(const uint8_t* h_stream, const uint8_t* d_stream, const std::vector<Chunk>& chunks,
                 const ChunkStream* cs, const SrcVol& dst, cudaStream_t st, bool is_2d);
  BatchBuffers b_;
  DecWork w_;
  size_t batch_first_ = 0;   // index of the current batch's first chunk in the caller's list
  bool whole_call_ = false;  // the current batch holds every chunk of the call
  rt::DBuf ids_, lis_off1_, tols_, obits_, ckey_[2], cval_[2], ccount_, coff_, csort_;
  rt::DBuf hdr_off_, hdr_cnt_, hdr_out_;   // chunk headers fetched from the device (no host copy given)
};

Compressor& shared_compressor();
Decompressor& shared_decompressor();
std::mutex& shared_api_mutex();

// Several GPUs behind one call of the host-pointer API (capi_multi.cu). multi_devices(): the devices
// SPERR_B200_DEVICES names (empty: one device, the calling thread's). The two functions return -2
// when the call cannot be split over them (the caller then takes the one-device path).
std::vector<int> multi_devices();
int comp_3d_multi(const void* src, int is_float, const size_t vol[3], const size_t chunk[3], int mode,
                  double quality, const std::vector<int>& devs, void** dst, size_t* dst_len);
// Volumes too large to sit in HBM beside the work buffers: one group of chunk slabs at a time
// (capi_stream.cu). -2: not needed / not possible, take the resident path.
int comp_3d_streamed(const void* src, int is_float, const size_t vol[3], const size_t chunk[3], int mode,
                     double quality, void** dst, size_t* dst_len);
int decomp_3d_streamed(const void* src, size_t src_len, int output_float, size_t* dimx, size_t* dimy,
                       size_t* dimz, void** dst);
int decomp_3d_multi(const void* src, size_t src_len, int output_float, const std::vector<int>& devs,
                    size_t* dimx, size_t* dimy, size_t* dimz, void** dst);

// Largest number of chunks processed at once (bounded by the list-key layout and by memory).
size_t pick_batch_chunks(const std::vector<Chunk>& chunks, size_t first, bool pwe);

}  // namespace sperr_b200
