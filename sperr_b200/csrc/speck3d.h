// SPECK3D integer coder: batch-level context and host drivers.
#pragma once

#include "kernels.h"

namespace sperr_b200 {

// Device pointers of the encoder's batch-wide work lists (passed to kernels by value).
struct EncCtx {
  ChunkDev* chunks;
  int nchunks;
  const ShapeDev* shapes;
  int maxp;

  // live sets of the current plane, sorted in list order, all chunks concatenated
  unsigned long long* rkey;
  node_t* rnode;
  unsigned* rseg;               // bits each root contributes to this plane's LIS part
  unsigned long long* rpos;     // exclusive scan of rseg (nroots + 1 entries)
  unsigned long long* rstart;   // first root of every chunk (nchunks + 1 entries)
  unsigned long long* total_roots;

  // sets that stay / become live for the next plane (unsorted)
  unsigned long long* ckey;
  node_t* cnode;
  unsigned long long* cand_count;
  unsigned long long cand_cap;

  // significant sets waiting to be expanded (ping-pong)
  node_t* fnode[2];
  unsigned long long* fpos[2];  // (absolute bit position << 10) | chunk
  unsigned long long* fcount;   // two counters
  unsigned long long front_cap;

  unsigned* err;

  // per (chunk, part, plane): size and absolute position of the LIP (part 0) / refinement (1) bits
  unsigned long long* sizes;
  unsigned long long* bases;
};

struct EncResult {
  int planes = 0;
  unsigned long long total_bits = 0;
  const uint32_t* payload = nullptr;  // device pointer to the bit array
  size_t payload_bytes = 0;           // ceil(min(budget, total_bits) / 8)
};

class Speck3DEncoder {
 public:
  // Expects for every chunk: mag, signs, pleaf filled; cmap set to -1; pyr_p / pyr_d allocated;
  // budget set (~0ull when unlimited). Chunks flagged is_const are skipped.
  void encode(ChunkDev* d_chunks, const std::vector<ChunkDev>& h_chunks,
              const ShapeDev* d_shapes, const std::vector<ShapeTables>& shapes,
              std::vector<EncResult>& results, cudaStream_t st);

 private:
  rt::DBuf ids_, keys_[2], nodes_[2], fnode_[2], fpos_[2], rseg_, rpos_, scan_tmp_, sort_tmp_,
      small_, stage_, counts_, sizes_;
};

struct DecResult {
  int ok = 1;
};

class Speck3DDecoder {
 public:
  // Expects for every chunk: mag / signs allocated; stream pointers set in the DecChunk array.
  void decode(ChunkDev* d_chunks, const std::vector<ChunkDev>& h_chunks, const ShapeDev* d_shapes,
              const std::vector<ShapeTables>& shapes, const uint8_t* const* d_streams,
              const std::vector<size_t>& stream_len, cudaStream_t st);

 private:
  rt::DBuf work_;
};

}  // namespace sperr_b200
