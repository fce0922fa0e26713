// SPECK3D integer decoder: tree policy of the 3D coder for the engine in speck_dec.cuh.
//   set partitioning      /root/reference/src/SPECK3D_INT.cpp:214-326 (m_partition_S_XYZ and friends)
//   initial sets          /root/reference/src/SPECK3D_INT.cpp:22-97
//   decoder significance  /root/reference/src/SPECK3D_INT_DEC.cpp:8-49
#include "speck_dec_fast.cuh"
#include "tree3d.cuh"

namespace sperr_b200 {

struct DecTree3D {
  struct Data {
    const ShapeDev* shapes;
  };

  static __device__ __forceinline__ int num_roots(const Data& t, const DecChunk& d, unsigned)
  {
    return t.shapes[d.shape].h->nroots;
  }

  static __device__ __forceinline__ void root(const Data& t, const DecChunk& d, unsigned, int r,
                                              node_t& nd, int& lis)
  {
    const RootDesc& rd = t.shapes[d.shape].h->roots[r];
    nd = make_node(rd.level, rd.ix, rd.iy, rd.iz);
    lis = rd.lis;
  }

  // Children in the reference's order (x fastest), empty ones dropped.
  static __device__ __forceinline__ int children(const Data& t, const DecChunk& d, unsigned,
                                                 node_t nd, int, DChild* out)
  {
    const ShapeDev s = t.shapes[d.shape];
    NodeGeom g;
    node_geom(s, nd, g);
    const ShapeHeader* h = s.h;
    int k = 0;
    if (g.Lc == h->leaf_level) {   // every child is a single coefficient
      for (unsigned cz = 0; cz < g.nzc; cz++)
        for (unsigned cy = 0; cy < g.nyc; cy++)
          for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
            DChild& r = out[k];
            r.pixel = 1;
            r.id = 0;
            r.lis = 0;
            r.idx = ((unsigned long long)(g.z0 + cz) * h->ny + (g.y0 + cy)) * h->nx + (g.x0 + cx);
          }
      return k;
    }
    for (unsigned cz = 0; cz < g.nzc; cz++)
      for (unsigned cy = 0; cy < g.nyc; cy++)
        for (unsigned cx = 0; cx < g.nxc; cx++, k++) {
          const unsigned jx = g.x0 + cx, jy = g.y0 + cy, jz = g.z0 + cz;
          DChild& r = out[k];
          unsigned lx, ly, lz;
          node_len(s, g.Lc, jx, jy, jz, lx, ly, lz);
          if (lx * ly * lz == 1) {
            r.pixel = 1;
            r.id = 0;
            r.lis = 0;
            r.idx = node_raster(s, g.Lc, jx, jy, jz);
          }
          else {
            r.pixel = 0;
            r.id = make_node(g.Lc, jx, jy, jz);
            r.idx = 0;
            r.lis = int(node_lis(s, g.Lc, jx, jy, jz));
          }
        }
    return k;
  }
};

void speck3d_decode(DecWork& w, const std::vector<DecJob>& jobs, const ShapeDev* d_shapes,
                    cudaStream_t st)
{
  DecTree3D::Data tree{d_shapes};
  run_decoder<DecTree3D>(w, jobs, tree, st);
}

}  // namespace sperr_b200
