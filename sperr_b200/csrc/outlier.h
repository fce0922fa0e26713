// PWE-mode outlier detection and SPECK1D outlier coding (see outlier.cu).
#pragma once

#include "speck.h"

namespace sperr_b200 {

class OutlierCoder {
 public:
  // Finds |orig - recon| > tol for every chunk (recon = ch.coef after the inverse transform,
  // orig = source value - mean), in ascending raster order. Fills `ooff` (per-chunk offsets into
  // the concatenated device arrays).
  void detect(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n, double tol,
              cudaStream_t st);
  // The same through an unordered append list that several kernels may feed (the fused inverse
  // transform records outliers while it rebuilds the values): begin_detect -> producers ->
  // end_detect. end_detect returns false when the list overflowed; the caller then repeats from
  // begin_detect (the list has been enlarged).
  OutlierSink begin_detect(int nchunks, size_t total_values, cudaStream_t st);
  void append_unfused(const SrcVol& src, const ChunkDev* d_chunks, int nchunks, size_t max_n, double tol,
                      const OutlierSink& sink, cudaStream_t st);
  bool end_detect(int nchunks, cudaStream_t st);
  // Alternative input for the stage-level test hook.
  void set_outliers(const std::vector<unsigned long long>& offsets, const unsigned* h_pos,
                    const double* h_err, cudaStream_t st);
  // SPECK1D-codes the outliers of every chunk. results[c].payload_bytes == 0 and planes == 0 for
  // chunks without outliers.
  void encode(const std::vector<unsigned long long>& total_len, double tol,
              std::vector<EncResult>& results, cudaStream_t st);

  std::vector<unsigned long long> ooff;  // nchunks + 1

 private:
  rt::DBuf cnt_, offs_, scan_tmp_, pick_, opos_, oerr_, meta_, omag_, osign_, nodes_, small_,
      pkeys_[2], pvals_[2], sort_tmp_, ppleaf_, pcmap_, pmag_, psigns_, first_, ochunks_, skey_[2],
      serr_[1], scount_;
  size_t sink_cap_ = 0;
  EncWork work_;
};

}  // namespace sperr_b200
