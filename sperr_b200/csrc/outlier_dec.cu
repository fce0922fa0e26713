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

}  // namespace sperr_b200
