// Work buffers of one batch of chunks, laid out as a few large HBM allocations that are
// sub-divided per chunk (grow-only, reused across batches and calls).
#pragma once

#include <map>

#include "kernels.h"

namespace sperr_b200 {

struct BatchBuffers {
  std::vector<ChunkDev> h;           // host mirror of the device chunk array
  std::vector<ShapeTables> shapes;   // distinct chunk shapes of this batch
  rt::DBuf d_chunks, d_shapes, shape_mem;
  rt::DBuf coef, mag, signs, pleaf, cmap, pyr_p, pyr_d, scratch;
  size_t max_n = 0;
  size_t sign_words = 0, mag_elems = 0;   // totals over the batch (padded)
  bool wide = false;

  static bool no_fused() { return std::getenv("SPERR_B200_NO_FUSED_DWT") != nullptr; }
  ChunkDev* dev() const { return d_chunks.as<ChunkDev>(); }
  const ShapeDev* dev_shapes() const { return d_shapes.as<ShapeDev>(); }
  int size() const { return int(h.size()); }

  // Lays out every per-chunk array. `need_coef`: fp64 buffer; `need_speck`: integer-coder arrays.
  // `need_enc`: also the encoder-only maps (msb positions, creation planes, significance pyramid).
  void setup(const std::vector<Chunk>& chunks, bool need_coef, bool need_speck, bool wide_mag,
             cudaStream_t st, bool need_enc = true, bool is_2d = false)
  {
    const int nc = int(chunks.size());
    h.assign(nc, ChunkDev());
    shapes.clear();
    std::map<std::array<uint32_t, 3>, int> seen;
    size_t tot_n = 0, tot_words = 0, tot_pyr = 0, tot_scr = 0;
    max_n = 0;
    wide = wide_mag;
    for (int c = 0; c < nc; c++) {
      const Chunk& k = chunks[c];
      ChunkDev& d = h[c];
      std::memset(&d, 0, sizeof(d));
      d.x0 = k.x0; d.y0 = k.y0; d.z0 = k.z0;
      d.nx = k.lx; d.ny = k.ly; d.nz = k.lz;
      d.n = k.nelem();
      const std::array<uint32_t, 3> key = {k.lx, k.ly, k.lz};
      auto it = seen.find(key);
      if (it == seen.end()) {
        it = seen.emplace(key, int(shapes.size())).first;
        shapes.push_back(build_shape(k.lx, k.ly, k.lz, is_2d));
      }
      d.shape = it->second;
      d.budget = ~0ull;
      d.min_key = ~0ull;
      d.max_key = 0;
      d.wide = wide_mag ? 1 : 0;
      {
        long long lo[8];
        d.fused = (need_coef && shapes[d.shape].h.dyadic >= 1 && !no_fused()) ? 1 : 0;
        if (d.fused)
          tot_scr += (fused_scratch_elems(k.lx, k.ly, k.lz, lo) + 63) & ~size_t(63);
      }
      tot_n += (d.n + 63) & ~size_t(63);
      tot_words += (d.n + 31) / 32 + 2;
      tot_pyr += shapes[d.shape].h.pyr_nodes + 64;
      max_n = std::max<size_t>(max_n, d.n);
    }
    if (need_coef) {
      coef.reserve(tot_n * 8);
      scratch.reserve(tot_scr * 8 + 64);
    }
    if (need_speck) {
      mag.reserve(tot_n * (wide_mag ? 8 : 4));
      signs.reserve(tot_words * 4);
      if (need_enc) {
        pleaf.reserve(tot_n);
        cmap.reserve(tot_n);
        pyr_p.reserve(tot_pyr);
        pyr_d.reserve(tot_pyr * 4);
      }
    }
    sign_words = tot_words;
    mag_elems = tot_n;
    size_t on = 0, ow = 0, op = 0, osc = 0;
    for (int c = 0; c < nc; c++) {
      ChunkDev& d = h[c];
      if (need_coef)
        d.coef = coef.as<double>() + on;
      if (d.fused) {
        long long lo[8];
        d.scratch = scratch.as<double>() + osc;
        osc += (fused_scratch_elems(d.nx, d.ny, d.nz, lo) + 63) & ~size_t(63);
      }
      if (need_speck) {
        d.mag = mag.as<unsigned char>() + on * (wide_mag ? 8 : 4);
        d.signs = signs.as<uint32_t>() + ow;
        if (need_enc) {
          d.pleaf = pleaf.as<int8_t>() + on;
          d.cmap = cmap.as<int8_t>() + on;
          d.pyr_p = pyr_p.as<int8_t>() + op;
          d.pyr_d = pyr_d.as<uint32_t>() + op;
        }
      }
      on += (d.n + 63) & ~size_t(63);
      ow += (d.n + 31) / 32 + 2;
      op += shapes[d.shape].h.pyr_nodes + 64;
    }
    if (need_speck && need_enc)
      rt::dset(cmap.p, 0xFF, tot_n, st);
    upload_shapes(st);
    d_chunks.reserve(sizeof(ChunkDev) * nc);
    push(st);
  }

  // Switch the magnitude array to 64-bit entries (fixed-rate high-precision retry, huge ranges).
  void make_wide(cudaStream_t st)
  {
    if (wide)
      return;
    size_t tot_n = 0;
    for (auto& d : h)
      tot_n += (d.n + 63) & ~size_t(63);
    mag.reserve(tot_n * 8);
    size_t on = 0;
    for (auto& d : h) {
      d.mag = mag.as<unsigned char>() + on * 8;
      on += (d.n + 63) & ~size_t(63);
    }
    wide = true;
    (void)st;
  }

  void push(cudaStream_t st) { rt::h2d(d_chunks.p, h.data(), sizeof(ChunkDev) * h.size(), st); }
  void pull(cudaStream_t st)
  {
    rt::d2h(h.data(), d_chunks.p, sizeof(ChunkDev) * h.size(), st);
    rt::sync(st);
  }

 private:
  void upload_shapes(cudaStream_t st)
  {
    // one allocation: [ShapeHeader | bnd | child0 | lev] per shape, 16-byte aligned pieces
    auto al = [](size_t v) { return (v + 15) & ~size_t(15); };
    size_t total = 0;
    for (auto& s : shapes)
      total += al(sizeof(ShapeHeader)) + al(s.bnd.size() * 4) + al(s.child0.size() * 4) + al(s.lev.size());
    shape_mem.reserve(total);
    std::vector<unsigned char> stage(total);
    std::vector<ShapeDev> sd(shapes.size());
    size_t off = 0;
    unsigned char* base = shape_mem.as<unsigned char>();
    for (size_t i = 0; i < shapes.size(); i++) {
      auto& s = shapes[i];
      std::memcpy(&stage[off], &s.h, sizeof(ShapeHeader));
      sd[i].h = reinterpret_cast<const ShapeHeader*>(base + off);
      off += al(sizeof(ShapeHeader));
      std::memcpy(&stage[off], s.bnd.data(), s.bnd.size() * 4);
      sd[i].bnd = reinterpret_cast<const uint32_t*>(base + off);
      off += al(s.bnd.size() * 4);
      std::memcpy(&stage[off], s.child0.data(), s.child0.size() * 4);
      sd[i].child0 = reinterpret_cast<const uint32_t*>(base + off);
      off += al(s.child0.size() * 4);
      std::memcpy(&stage[off], s.lev.data(), s.lev.size());
      sd[i].lev = reinterpret_cast<const uint8_t*>(base + off);
      off += al(s.lev.size());
    }
    rt::h2d(shape_mem.p, stage.data(), total, st);
    d_shapes.reserve(sizeof(ShapeDev) * sd.size());
    rt::h2d(d_shapes.p, sd.data(), sizeof(ShapeDev) * sd.size(), st);
    rt::sync(st);  // `stage` and `sd` are about to go out of scope
  }
};

}  // namespace sperr_b200
