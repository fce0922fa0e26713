#include "pipeline.h"

#include <cmath>
#include <condition_variable>
#include <exception>
#include <limits>
#include <thread>

namespace sperr_b200 {

namespace {

// inverse of the order-preserving key used for atomic min / max of doubles
double key_to_double(unsigned long long k)
{
  const unsigned long long b = (k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k;
  double v;
  std::memcpy(&v, &b, 8);
  return v;
}

struct Piece {
  const unsigned char* src;
  unsigned long long dst_off;
  unsigned long long len;
};

}  // namespace

// One pipeline pair and one lock per DEVICE: the host-pointer API, the device-pointer API and the
// chunk-range API all reuse the same grow-only work buffers (tens of GB for a 1024^3 call) of the
// device the calling thread has selected.
Compressor& shared_compressor()
{
  static Compressor* p[rt::kMaxDevices] = {};
  static std::mutex mu;
  std::lock_guard<std::mutex> l(mu);
  Compressor*& c = p[rt::cur_dev()];
  if (!c)
    c = new Compressor();
  return *c;
}
Decompressor& shared_decompressor()
{
  static Decompressor* p[rt::kMaxDevices] = {};
  static std::mutex mu;
  std::lock_guard<std::mutex> l(mu);
  Decompressor*& c = p[rt::cur_dev()];
  if (!c)
    c = new Decompressor();
  return *c;
}
// One job at a time per DEVICE: the work buffers of a device are shared by all entry points.
std::mutex& shared_api_mutex()
{
  static std::mutex* m = new std::mutex[rt::kMaxDevices];
  return m[rt::cur_dev()];
}

// One block column per piece; bytes are copied with a grid-stride loop.
__global__ void k_copy_pieces(const Piece* pieces, unsigned char* dst)
{
  const Piece p = pieces[blockIdx.y];
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < p.len;
       i += (unsigned long long)gridDim.x * blockDim.x)
    dst[p.dst_off + i] = p.src[i];
}

size_t pick_batch_chunks(const std::vector<Chunk>& chunks, size_t first, bool pwe)
{
  // Rough HBM footprint per value: fp64 buffer 8, magnitudes 4 (8 when wide), msb / creation maps
  // 2, pyramid ~0.8, payload staging ~2-4, list buffers ~9 => budget 40 bytes per value.
  size_t free_b = size_t(64) << 30, total_b = 0;
#ifndef SPERR_EMUL
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess)
    free_b = size_t(64) << 30;
  free_b += rt::DBuf::held();   // our own grow-only buffers are reused, not allocated again
#endif
  (void)total_b;
  (void)pwe;
  const double budget = double(free_b) * 0.7;
  double used = 0;
  size_t n = 0;
  while (first + n < chunks.size() && n < 256) {
    const double need = double(chunks[first + n].nelem()) * 40.0 + (64 << 20);
    if (n > 0 && used + need > budget)
      break;
    used += need;
    n++;
  }
  return std::max<size_t>(n, 1);
}

void Compressor::compress(const SrcVol& src, const std::vector<Chunk>& chunks, int mode,
                          double quality, bool is_2d, rt::DBuf& d_out, std::vector<size_t>& lens,
                          cudaStream_t st)
{
  lens.assign(chunks.size(), 0);
  size_t used = 0;  // bytes of d_out in use
  size_t first = 0;
  while (first < chunks.size()) {
    size_t nb = pick_batch_chunks(chunks, first, mode == kModePWE);
    if (max_batch)
      nb = std::min(nb, max_batch);
    if (before_batch)
      before_batch(first, nb);
    batch_first_ = first;
    whole_call_ = nb == chunks.size();
    std::vector<Chunk> sub(chunks.begin() + first, chunks.begin() + first + nb);
    std::vector<std::vector<uint8_t>> hdrs;
    run_batch(src, sub, mode, quality, is_2d, hdrs, st);

    // ---- assemble: condi(17) | speck header(9) | payload | [outlier header(9) | payload] ----
    auto has_outliers = [&](size_t c) {
      return mode == kModePWE && !b_.h[c].is_const && out_.ooff[c + 1] > out_.ooff[c];
    };
    std::vector<Piece> pieces;
    std::vector<unsigned char> hbytes;
    struct Ref { size_t hoff, hlen; };
    size_t batch_total = 0;
    for (size_t c = 0; c < nb; c++) {
      size_t len = hdrs[c].size();
      if (!b_.h[c].is_const) {
        len += spk_res_[c].payload_bytes;
        if (has_outliers(c))
          len += 9 + out_res_[c].payload_bytes;
      }
      lens[first + c] = len;
      batch_total += len;
    }
    if (used + batch_total > d_out.bytes) {
      rt::DBuf bigger(std::max(used + batch_total, d_out.bytes * 2));
      rt::d2d(bigger.p, d_out.p, used, st);
      rt::sync(st);
      d_out = std::move(bigger);
    }
    size_t off = used;
    std::vector<size_t> hdr_off(nb), ohdr_off(nb, 0);
    for (size_t c = 0; c < nb; c++) {
      hdr_off[c] = hbytes.size();
      hbytes.insert(hbytes.end(), hdrs[c].begin(), hdrs[c].end());
      if (!b_.h[c].is_const && has_outliers(c)) {
        ohdr_off[c] = hbytes.size();
        unsigned char oh[9];
        oh[0] = (unsigned char)out_res_[c].planes;
        const unsigned long long tb = out_res_[c].total_bits;
        std::memcpy(oh + 1, &tb, 8);
        hbytes.insert(hbytes.end(), oh, oh + 9);
      }
    }
    asm_hdr_.reserve(hbytes.size() + 16);   // grow-only members: no cudaMalloc / cudaFree per call
    rt::h2d(asm_hdr_.p, hbytes.data(), hbytes.size(), st);
    const unsigned char* dh = asm_hdr_.as<unsigned char>();
    for (size_t c = 0; c < nb; c++) {
      pieces.push_back({dh + hdr_off[c], off, hdrs[c].size()});
      off += hdrs[c].size();
      if (b_.h[c].is_const)
        continue;
      if (spk_res_[c].payload_bytes) {
        pieces.push_back({reinterpret_cast<const unsigned char*>(spk_res_[c].payload), off,
                          spk_res_[c].payload_bytes});
        off += spk_res_[c].payload_bytes;
      }
      if (has_outliers(c)) {
        pieces.push_back({dh + ohdr_off[c], off, 9});
        off += 9;
        if (out_res_[c].payload_bytes) {
          pieces.push_back({reinterpret_cast<const unsigned char*>(out_res_[c].payload), off,
                            out_res_[c].payload_bytes});
          off += out_res_[c].payload_bytes;
        }
      }
    }
    asm_pieces_.reserve(pieces.size() * sizeof(Piece));
    rt::h2d(asm_pieces_.p, pieces.data(), pieces.size() * sizeof(Piece), st);
    const Piece* dp = asm_pieces_.as<Piece>();
    unsigned char* dst = d_out.as<unsigned char>();
    LAUNCH(k_copy_pieces, dim3(64, unsigned(pieces.size())), dim3(256), 0, st, dp, dst);
    rt::sync(st);
    used = off;
    first += nb;
  }
}

void Compressor::run_batch(const SrcVol& src, const std::vector<Chunk>& chunks, int mode,
                           double quality, bool is_2d, std::vector<std::vector<uint8_t>>& hdrs,
                           cudaStream_t st)
{
  const int nc = int(chunks.size());
  b_.setup(chunks, true, true, false, st, true, is_2d);

  // ---- conditioner ----
  std::vector<unsigned> ns(nc);
  int max_strides = 1;
  for (int c = 0; c < nc; c++) {
    ns[c] = unsigned(mean_num_strides(chunks[c].nelem()));
    max_strides = std::max<int>(max_strides, int(ns[c]));
  }
  stride_mean_.reserve((size_t)nc * max_strides * 8);
  nstrides_.reserve(nc * 4);
  not_const_.reserve(nc * 4);
  rt::h2d(nstrides_.p, ns.data(), nc * 4, st);
  rt::dset(not_const_.p, 0, nc * 4, st);
  bool any_unfused = false, any_fused = false;
  size_t total_values = 0;
  for (auto& d : b_.h) {
    (d.fused ? any_fused : any_unfused) = true;
    total_values += d.n;
  }
  // Front end by slab groups (host-pointer API with a pinned source): the statistics and the forward
  // transform of a group start as soon as `before_group` says its part of the volume has arrived,
  // while the rest is still on its way over PCIe. Needs every chunk on the fused transform path.
  const bool by_groups = before_group && group_chunks > 0 && !any_unfused && size_t(nc) > group_chunks &&
                         whole_call_;
  if (!by_groups) {
    if (before_group)
      before_group(batch_first_, size_t(nc));
    rt::ProfScope ps("c.stats", st);
    launch_stats(src, b_.dev(), nc, stride_mean_.as<double>(), max_strides, nstrides_.as<unsigned>(),
                 not_const_.as<unsigned>(), mode == kModePSNR, st);
  }
  if (any_unfused) {
    rt::ProfScope ps("c.gather", st);
    launch_gather(src, b_.dev(), nc, b_.max_n, st);
  }

  // ---- wavelet transform, one launch series per distinct chunk shape ----
  std::vector<std::vector<int>> groups(b_.shapes.size());
  for (int c = 0; c < nc; c++)
    groups[b_.h[c].shape].push_back(c);
  std::vector<int> flat;
  std::vector<size_t> goff;
  for (auto& g : groups) {
    goff.push_back(flat.size());
    flat.insert(flat.end(), g.begin(), g.end());
  }
  ids_.reserve(flat.size() * 4 + 4);
  rt::h2d(ids_.p, flat.data(), flat.size() * 4, st);
  rt::sync(st);
  // dyadic shapes: fused kernels (dwt_fused.cu) that read the volume themselves and track the
  // coefficient maximum; everything else: gather, per-axis passes in place, separate maximum
  auto group_fused = [&](size_t s) { return b_.h[groups[s][0]].fused != 0; };
  auto transform = [&](bool inverse, bool fused_groups, const OutlierSink& sink, cudaStream_t st) {
    rt::ProfScope ps(inverse ? "c.idwt" : "c.dwt", st);
    for (size_t s = 0; s < groups.size(); s++) {
      if (groups[s].empty() || group_fused(s) != fused_groups)
        continue;
      const ShapeHeader& h = b_.shapes[s].h;
      const int* ids = ids_.as<int>() + goff[s];
      const int n = int(groups[s].size());
      if (!fused_groups)
        launch_dwt(inverse, b_.dev(), ids, n, h.nx, h.ny, h.nz, is_2d, st);
      else if (!inverse)
        launch_dwt_fused_forward(src, b_.dev(), ids, n, h.nx, h.ny, h.nz, st);
      else   // PWE: rebuild the values, compare with the source, record the outliers
        launch_dwt_fused_inverse(src, 2, b_.dev(), ids, n, h.nx, h.ny, h.nz, quality, sink,
                                 CorrectorList{nullptr, nullptr, nullptr}, st);
    }
  };
  if (by_groups) {
    for (size_t g0 = 0; g0 < size_t(nc); g0 += group_chunks) {
      const size_t g1 = std::min(size_t(nc), g0 + group_chunks);
      before_group(batch_first_ + g0, g1 - g0);
      {
        rt::ProfScope ps("c.stats", st);
        launch_stats(src, b_.dev() + g0, int(g1 - g0), stride_mean_.as<double>() + g0 * size_t(max_strides),
                     max_strides, nstrides_.as<unsigned>() + g0, not_const_.as<unsigned>() + g0,
                     mode == kModePSNR, st);
      }
      rt::ProfScope ps("c.dwt", st);
      for (size_t si = 0; si < groups.size(); si++) {
        if (groups[si].empty())
          continue;
        // chunk ids of a shape group ascend: the members inside [g0, g1) are one sub-range
        const auto lo = std::lower_bound(groups[si].begin(), groups[si].end(), int(g0));
        const auto hi = std::lower_bound(groups[si].begin(), groups[si].end(), int(g1));
        if (lo == hi)
          continue;
        const ShapeHeader& h = b_.shapes[si].h;
        launch_dwt_fused_forward(src, b_.dev(), ids_.as<int>() + goff[si] + (lo - groups[si].begin()),
                                 int(hi - lo), h.nx, h.ny, h.nz, st);
      }
    }
  }
  else
    transform(false, true, OutlierSink{}, st);
  if (any_unfused) {
    transform(false, false, OutlierSink{}, st);
    rt::ProfScope ps("c.absmax", st);
    launch_absmax(b_.dev(), nc, b_.max_n, st);
  }
  b_.pull(st);

  // ---- quantisation step per chunk (SPECK_FLT::m_estimate_q, src/SPECK_FLT.cpp:268-309) ----
  std::vector<double> maxabs(nc);
  for (int c = 0; c < nc; c++) {
    ChunkDev& d = b_.h[c];
    std::memcpy(&maxabs[c], &d.max_bits, 8);
    if (d.is_const)
      continue;
    if (mode == kModePWE)
      d.q = quality * 1.5;
    else if (mode == kModeRate) {
      d.q = maxabs[c] / double(std::numeric_limits<uint32_t>::max());
      size_t bud = size_t(quality * double(d.n));
      while (bud % 8 != 0)
        bud++;
      d.budget = bud == 0 ? ~0ull : bud;
    }
  }
  if (mode == kModePSNR) {
    std::vector<double> t_mse(nc, 0.0), q(nc, 0.0);
    std::vector<int> todo;
    size_t max_s = 1;
    for (int c = 0; c < nc; c++) {
      ChunkDev& d = b_.h[c];
      if (d.is_const)
        continue;
      const double mx = key_to_double(d.max_key) - d.mean;
      const double mn = key_to_double(d.min_key) - d.mean;
      const double range = mx - mn;
      t_mse[c] = (range * range) * std::pow(10.0, -quality / 10.0);
      q[c] = 2.0 * std::sqrt(t_mse[c] * 3.0);
      todo.push_back(c);
      max_s = std::max<size_t>(max_s, d.n / 4096 + 1);
    }
    mse_ids_.reserve(nc * 4);
    mse_q_.reserve(nc * 8);
    mse_out_.reserve(nc * 8);
    mse_part_.reserve((size_t)nc * max_s * 8);
    while (!todo.empty()) {
      std::vector<double> qs;
      for (int c : todo)
        qs.push_back(q[c]);
      rt::h2d(mse_ids_.p, todo.data(), todo.size() * 4, st);
      rt::h2d(mse_q_.p, qs.data(), qs.size() * 8, st);
      launch_mse(b_.dev(), mse_ids_.as<int>(), mse_q_.as<double>(), mse_part_.as<double>(), int(max_s),
                 mse_out_.as<double>(), int(todo.size()), st);
      std::vector<double> mse(todo.size());
      rt::d2h(mse.data(), mse_out_.p, todo.size() * 8, st);
      rt::sync(st);
      std::vector<int> next;
      for (size_t i = 0; i < todo.size(); i++) {
        const int c = todo[i];
        // NaN compares false, exactly as in the reference's while loop
        if (mse[i] > t_mse[c]) {
          q[c] /= std::exp2(0.25);
          next.push_back(c);
        }
      }
      todo.swap(next);
    }
    for (int c = 0; c < nc; c++)
      if (!b_.h[c].is_const)
        b_.h[c].q = q[c];
  }
  b_.push(st);

  auto quantize_and_encode = [&](Speck3DEncoder& enc, std::vector<EncResult>& res) {
    // counters read back by this thread and by the outlier thread below go through pinned bounce
    // buffers, so that neither thread's read-back holds up the other's launches (rt.h)
    rt::readback_abandon();   // nothing of an earlier, failed call may be delivered now
    rt::ReadbackScope readback_main;
    (void)readback_main;
    launch_qdecide(b_.dev(), nc, st);
    b_.pull(st);
    bool any_wide = false;
    for (auto& d : b_.h) {
      if (!d.is_const && d.fe_invalid)
        throw std::runtime_error("FE_INVALID while quantising");
      any_wide |= (!d.is_const && d.wide);
    }
    if (any_wide && !b_.wide) {
      b_.make_wide(st);
      b_.push(st);
    }
    {
      rt::ProfScope ps("c.quantize", st);
      launch_quantize(b_.dev(), nc, b_.max_n, st);
    }
    // PWE: the outlier path (de-quantise, inverse transform, compare with the source, SPECK1D-code
    // the differences) only needs the quantised integers, like the SPECK3D encoder, and both are
    // chains of small launches with host round trips: run them side by side on two streams.
    auto outlier_chain = [&](cudaStream_t s) {
      {
        rt::ProfScope ps("c.inv_quantize", s);
        launch_inv_quantize(b_.dev(), nc, b_.max_n, s);
      }
      if (any_unfused)
        transform(true, false, OutlierSink{}, s);
      for (;;) {
        const OutlierSink sink = out_.begin_detect(nc, total_values, s);
        if (any_fused)
          transform(true, true, sink, s);
        if (any_unfused) {
          rt::ProfScope ps("c.outlier_detect", s);
          out_.append_unfused(src, b_.dev(), nc, b_.max_n, quality, sink, s);
        }
        if (out_.end_detect(nc, s))
          break;
      }
      std::vector<unsigned long long> tl(nc);
      for (int c = 0; c < nc; c++)
        tl[c] = b_.h[c].n;
      rt::ProfScope ps("c.outlier_encode", s);
      out_.encode(tl, quality, out_res_, s);
    };
    cudaStream_t enc_st = st;
    auto speck_chain = [&] {
      rt::ProfScope ps("c.speck3d", enc_st);
      enc.encode(b_.dev(), b_.h, b_.dev_shapes(), b_.shapes, res, enc_st);
    };
    if (mode != kModePWE) {
      speck_chain();
      return;
    }
#ifdef SPERR_EMUL
    outlier_chain(st);
    speck_chain();
#else
    if (!side_) {
      RT_CHECK(cudaStreamCreateWithFlags(&side_, cudaStreamNonBlocking));
      RT_CHECK(cudaEventCreateWithFlags(&side_ev_, cudaEventDisableTiming));
    }
    RT_CHECK(cudaEventRecord(side_ev_, st));
    RT_CHECK(cudaStreamWaitEvent(side_, side_ev_, 0));
    // Both chains share the GPU. The encoder is the longer one and mostly small, latency-bound
    // launches; the outlier chain starts with two bandwidth kernels that fill every SM (measured:
    // enc.lipref_count 15 ms beside k_inv3d<2>, 3.3 ms alone). With the encoder on a stream of the
    // highest priority the block scheduler hands freed CTA slots to its kernels first (the inverse
    // transform of the outlier chain runs in short z segments for that reason, dwt_fused.cu).
    if (std::getenv("SPERR_B200_PRIO")) {   // measured: no gain, off by default
      if (!hi_) {
        int least = 0, greatest = 0;
        RT_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest));
        RT_CHECK(cudaStreamCreateWithPriority(&hi_, cudaStreamNonBlocking, greatest));
        RT_CHECK(cudaEventCreateWithFlags(&hi_ev_, cudaEventDisableTiming));
      }
      RT_CHECK(cudaStreamWaitEvent(hi_, side_ev_, 0));
      enc_st = hi_;
    }
    int dev = 0;
    RT_CHECK(cudaGetDevice(&dev));
    // Experiment (SPERR_B200_STAGGER=1, off by default): both chains start with their
    // bandwidth-bound kernels (pyramid + LIP / refinement counts here, de-quantisation + inverse
    // transform there: enc.lipref_count takes 14.9 ms beside k_inv3d<2> against 3.3 ms alone) and
    // end with latency-bound loops that leave the GPU idle. Staggered, the outlier chain starts when
    // the encoder enters its bit-plane loop, so that its heavy kernels fill that loop's idle GPU.
    struct Gate {
      std::mutex m;
      std::condition_variable cv;
      bool open = false;
      void release()
      {
        {
          std::lock_guard<std::mutex> l(m);
          open = true;
        }
        cv.notify_all();
      }
      void wait()
      {
        std::unique_lock<std::mutex> l(m);
        cv.wait(l, [this] { return open; });
      }
    } gate;
    // Where the outlier chain is let go (SPERR_B200_GATE): "none" (default) at once; "setup" when the
    // encoder has queued the zeroing of its staging arrays and tables; "loop" (the SPERR_B200_STAGGER
    // experiment) when it enters its bit-plane loop. Measured on B200 (1024^3): compress 53.5 / 53.2 /
    // 53.2 ms, c.speck3d 40.1 / 40.8 / 40.7 ms -- the gate only moves the contention (none:
    // enc.stage_zero 6.9, pyramid 5.9, plane loop 16.4 ms; setup: 0.3, 3.0, 23.9; loop: 0.3, 3.0,
    // 26.6): both chains need the whole GPU for about as long as they run, and the sum is what it is.
    const char* const gate_env = std::getenv("SPERR_B200_GATE");
    const std::string gate_at = std::getenv("SPERR_B200_STAGGER") ? "loop" : (gate_env ? gate_env : "none");
    const bool stagger = gate_at == "loop" || gate_at == "setup";
    if (stagger) {
      if (!stagger_ev_)
        RT_CHECK(cudaEventCreateWithFlags(&stagger_ev_, cudaEventDisableTiming));
      auto open_gate = [&](cudaStream_t s) {
        RT_CHECK(cudaEventRecord(stagger_ev_, s));
        gate.release();
      };
      if (gate_at == "loop")
        enc.set_before_plane_loop(open_gate);
      else
        enc.set_after_setup(open_gate);
    }
    else
      gate.release();
    std::exception_ptr side_err;
    std::thread helper([&] {
      try {
        rt::ReadbackScope readback_side;
        (void)readback_side;
        RT_CHECK(cudaSetDevice(dev));
        gate.wait();
        if (stagger)
          RT_CHECK(cudaStreamWaitEvent(side_, stagger_ev_, 0));
        outlier_chain(side_);
        rt::sync(side_);
      }
      catch (...) {
        side_err = std::current_exception();
      }
    });
    try {
      speck_chain();
    }
    catch (...) {
      enc.set_before_plane_loop(nullptr);
      enc.set_after_setup(nullptr);
      gate.release();   // the encoder may have failed before it reached its plane loop
      helper.join();
      throw;
    }
    enc.set_before_plane_loop(nullptr);
    enc.set_after_setup(nullptr);
    gate.release();   // (no-op unless the encoder had nothing to code)
    helper.join();
    if (enc_st != st) {   // (the encoder has synchronised its stream; this orders later work on st after it)
      RT_CHECK(cudaEventRecord(hi_ev_, enc_st));
      RT_CHECK(cudaStreamWaitEvent(st, hi_ev_, 0));
    }
    if (side_err)
      std::rethrow_exception(side_err);
#endif
  };
  quantize_and_encode(enc_, spk_res_);

  // ---- fixed-rate: not enough bits produced => redo with the high-precision step (:530-538) ----
  if (mode == kModeRate) {
    std::vector<int> retry;
    for (int c = 0; c < nc; c++) {
      if (b_.h[c].is_const)
        continue;
      const unsigned long long actual = (9ull + spk_res_[c].payload_bytes) * 8ull;
      const size_t budget = size_t(quality * double(b_.h[c].n));
      if (actual < budget)
        retry.push_back(c);
    }
    if (!retry.empty()) {
      // keep the finished chunks out of the second pass by flagging them constant for a moment
      std::vector<int> saved(nc);
      std::vector<double> q_keep(nc);
      for (int c = 0; c < nc; c++) {
        saved[c] = b_.h[c].is_const;
        q_keep[c] = b_.h[c].q;
        b_.h[c].is_const = 1;
      }
      for (int c : retry) {
        b_.h[c].is_const = 0;
        b_.h[c].q = maxabs[c] / 0x1.fffffffffffffp52;
        b_.h[c].wide = 0;
        b_.h[c].fe_invalid = 0;
      }
      b_.push(st);
      rt::dset(b_.cmap.p, 0xFF, b_.cmap.bytes, st);
      std::vector<EncResult> res2;
      // the second pass needs its own payload staging: the first-pass payloads are still needed
      quantize_and_encode(enc_hp_, res2);
      for (int c : retry)
        spk_res_[c] = res2[c];
      for (int c = 0; c < nc; c++) {
        const bool was_retry = std::find(retry.begin(), retry.end(), c) != retry.end();
        b_.h[c].is_const = saved[c];
        if (!was_retry)
          b_.h[c].q = q_keep[c];
      }
    }
  }
  else {
    b_.pull(st);
  }

  // ---- per-chunk headers: conditioner (17 bytes) + SPECK stream header (9 bytes) ----
  hdrs.assign(nc, {});
  for (int c = 0; c < nc; c++) {
    const ChunkDev& d = b_.h[c];
    std::vector<uint8_t>& h = hdrs[c];
    h.assign(17, 0);
    if (d.is_const) {
      // Conditioner::condition, src/Conditioner.cpp:28-44: {0x81, u64 nval, f64 value}
      h[0] = 0x81;
      const uint64_t nval = d.n;
      std::memcpy(&h[1], &nval, 8);
      std::memcpy(&h[9], &d.first_val, 8);
      continue;
    }
    h[0] = 0x80;
    std::memcpy(&h[1], &d.mean, 8);
    std::memcpy(&h[9], &d.q, 8);  // Conditioner::save_q, :104-108
    h.resize(26);
    h[17] = uint8_t(spk_res_[c].planes);
    const uint64_t tb = spk_res_[c].total_bits;
    std::memcpy(&h[18], &tb, 8);
  }
}

}  // namespace sperr_b200
