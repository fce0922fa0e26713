// Integer SPECK coders: batch-level context and host drivers.
#pragma once

#include <vector>
#include <utility>
#include <functional>

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
  unsigned long long* fcount;   // three rotating counters (k_lis_plane)
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

// Work buffers of the bit-plane loop (grow-only, reused across calls).
struct EncWork {
  rt::DBuf keys[2], nodes[2], fnode[2], fpos[2], rseg, rpos, scan_tmp, sort_tmp, small, stage,
      counts, sizes;
  // called once per run, right after the last bandwidth-bound launch before the bit-plane loop
  // (pyramid and LIP / refinement counts are queued; what follows is latency-bound): the point at
  // which work of another stream no longer competes with this encoder for the memory system
  std::function<void(cudaStream_t)> before_plane_loop;
  // called once per run when the zeroing of the staging arrays and count tables has been queued (a
  // string of small memory operations: each of them waits for a CTA slot if another stream is
  // flooding the GPU meanwhile)
  std::function<void(cudaStream_t)> after_setup;
  // `stage` (the zeroed bit arrays the coder ORs its output into) is sized for the worst case, 69 MB
  // per 256^3 chunk; a run writes a few MB of it. It is zeroed as a whole only when it is new (or a
  // run ended in an exception); otherwise a run clears just what the run before wrote
  // (stage_dirty: word offset, words).
  bool stage_clean = false;
  std::vector<std::pair<size_t, size_t>> stage_dirty;
};

class Speck3DEncoder {
 public:
  // Expects for every chunk: mag, signs, pleaf filled; cmap set to -1; pyr_p / pyr_d allocated;
  // budget set (~0ull when unlimited). Chunks flagged is_const are skipped.
  void encode(ChunkDev* d_chunks, const std::vector<ChunkDev>& h_chunks,
              const ShapeDev* d_shapes, const std::vector<ShapeTables>& shapes,
              std::vector<EncResult>& results, cudaStream_t st);
  void set_before_plane_loop(std::function<void(cudaStream_t)> f) { work_.before_plane_loop = std::move(f); }
  void set_after_setup(std::function<void(cudaStream_t)> f) { work_.after_setup = std::move(f); }

 private:
  rt::DBuf ids_;
  EncWork work_;
};

}  // namespace sperr_b200
