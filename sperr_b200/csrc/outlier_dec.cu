// Outlier decoding for PWE-mode streams: SPECK1D decode of the sparse corrector array and its
// application to the reconstructed values.
//   SPECK1D_INT / _DEC          /root/reference/src/SPECK1D_INT.cpp:18-56, src/SPECK1D_INT_DEC.cpp:12-125
//   Outlier_Coder::decode       /root/reference/src/Outlier_Coder.cpp:133-149, m_inverse_quantize :206-234
//   application                 /root/reference/src/SPECK_FLT.cpp:576-585
#include "speck_dec_fast.cuh"

namespace sperr_b200 {

// node = start | len << 32 ; list index = depth of the set (the two initial halves are depth 1)
struct DecTree1D {
  struct Data {
    int unused;
  };
  static __device__ __forceinline__ int num_roots(const Data&, const DecChunk&, unsigned) { return 2; }
  static __device__ __forceinline__ void root(const Data&, const DecChunk& d, unsigned, int r,
                                              node_t& nd, int& lis)
  {
    const unsigned long long n = d.n, first = n - n / 2;
    nd = r == 0 ? (first << 32) : (first | ((n / 2) << 32));
    lis = 1;
  }
  static __device__ __forceinline__ int children(const Data&, const DecChunk&, unsigned, node_t nd,
                                                 int lis, DChild* out)
  {
    const unsigned long long start = nd & 0xffffffffull, len = nd >> 32;
    const unsigned long long l0 = len - len / 2, l1 = len / 2;
    out[0].pixel = l0 == 1;
    out[0].idx = start;
    out[0].id = start | (l0 << 32);
    out[0].lis = lis + 1;
    out[1].pixel = l1 == 1;
    out[1].idx = start + l0;
    out[1].id = (start + l0) | (l1 << 32);
    out[1].lis = lis + 1;
    return 2;
  }
};

void speck1d_decode(DecWork& w, const std::vector<DecJob>& jobs, cudaStream_t st)
{
  DecTree1D::Data tree{0};
  run_decoder<DecTree1D>(w, jobs, tree, st);
}

// coef[i] += corrector for every decoded outlier (the significant pixels of the 1D decode)
__global__ void k_outlier_apply(const DecChunk* jobs, const ChunkDev* chunks, const double* tols)
{
  const unsigned c = blockIdx.y;
  const DecChunk& d = jobs[c];
  if (d.skip || d.planes == 0)
    return;
  const ChunkDev& ch = chunks[c];
  const double tol = tols[c];
  const unsigned long long words = (d.n + 31) / 32;
  for (unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; j < words;
       j += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned m = d.lsp[j];
    while (m) {
      const int bit = __ffs(m) - 1;
      m &= m - 1;
      const unsigned long long i = j * 32 + bit;
      const unsigned long long mg = d.wide ? reinterpret_cast<const unsigned long long*>(d.mag)[i]
                                           : (unsigned long long)reinterpret_cast<const unsigned*>(d.mag)[i];
      if (mg == 0)
        continue;
      const bool pos = (d.signs[i >> 5] >> (i & 31)) & 1u;
      double e = mg == 1 ? 1.1 : __dsub_rn(__ull2double_rn(mg), 0.25);
      e = __dmul_rn(e, __dmul_rn(tol, pos ? 1.0 : -1.0));
      ch.coef[i] = __dadd_rn(ch.coef[i], e);
    }
  }
}

void launch_outlier_apply(const DecChunk* d_jobs, const ChunkDev* d_chunks, const double* d_tols,
                          int nchunks, size_t max_n, cudaStream_t st)
{
  const unsigned gx = unsigned(std::min<size_t>((max_n / 32 + 255) / 256 + 1, 512));
  LAUNCH(k_outlier_apply, dim3(gx, nchunks), dim3(256), 0, st, d_jobs, d_chunks, d_tols);
}

}  // namespace sperr_b200
