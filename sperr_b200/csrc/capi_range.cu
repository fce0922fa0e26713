// Chunk-range entry points for one-process-per-GPU jobs (include/sperr_b200.h section 2b).
//
// SPERR's chunks are coded independently (/root/reference/src/SPERR3D_OMP_C.cpp:94-130,
// src/SPERR3D_OMP_D.cpp:94-130), so a volume shards along chunk_volume's order with no halo: rank r
// codes chunks [begin, end) out of the bounding box of those chunks, which it holds in its own
// device memory. The only exchange is of the chunk lengths and the compressed bytes (the caller's
// collective: torch.distributed over NCCL in sperr_b200/sharded.py); the container header is
// assembled here so that the result is the reference's single-stream layout
// (SPERR3D_OMP_C::m_generate_header, src/SPERR3D_OMP_C.cpp:163-234).
#include "../../include/sperr_b200.h"

#include <mutex>

#include "pipeline.h"

using namespace sperr_b200;

namespace {

struct RangeState {   // per device, like the pipelines it uses
  Compressor* comp = nullptr;
  Decompressor* decomp = nullptr;
  rt::DBuf out, stream;
};
RangeState& range_state()
{
  static RangeState s[rt::kMaxDevices];
  return s[rt::cur_dev()];
}
#define g_rmutex (shared_api_mutex())
#define g_rcomp (range_state().comp)
#define g_rdecomp (range_state().decomp)
#define g_rout (range_state().out)
#define g_rstream (range_state().stream)

template <typename F>
int guarded(F&& f)
{
  try {
    return f();
  }
  catch (const std::exception& e) {
    if (std::getenv("SPERR_B200_VERBOSE"))
      std::fprintf(stderr, "sperr_b200: %s\n", e.what());
    return -1;
  }
}

bool range_chunks(const size_t vol[3], const size_t chunk[3], size_t begin, size_t end,
                  const size_t origin[3], const size_t extent[3], std::vector<Chunk>& out)
{
  size_t cd[3];
  for (int i = 0; i < 3; i++) {
    if (vol[i] == 0 || vol[i] > 0xFFFFFFFFull)
      return false;
    cd[i] = std::min(std::max<size_t>(1, chunk[i]), vol[i]);
  }
  const auto all = chunk_volume(vol, cd);
  if (begin > end || end > all.size())
    return false;
  out.assign(all.begin() + begin, all.begin() + end);
  for (auto& c : out) {
    if (c.x0 < origin[0] || c.y0 < origin[1] || c.z0 < origin[2] ||
        c.x0 + c.lx > origin[0] + extent[0] || c.y0 + c.ly > origin[1] + extent[1] ||
        c.z0 + c.lz > origin[2] + extent[2])
      return false;   // chunk outside the caller's box
    if (c.lx > 65535 || c.ly > 65535 || c.lz > 65535 || c.nelem() >= (1ull << 31))
      return false;
    c.x0 -= uint32_t(origin[0]);
    c.y0 -= uint32_t(origin[1]);
    c.z0 -= uint32_t(origin[2]);
  }
  return true;
}

}  // namespace

extern "C" {

size_t sperr_b200_num_chunks(const size_t vol[3], const size_t chunk[3])
{
  size_t cd[3];
  for (int i = 0; i < 3; i++) {
    if (vol[i] == 0)
      return 0;
    cd[i] = std::min(std::max<size_t>(1, chunk[i]), vol[i]);
  }
  size_t nseg[3], n = 0;
  return chunk_grid(vol, cd, nseg, &n) ? n : 0;
}

int sperr_b200_chunk_box(const size_t vol[3], const size_t chunk[3], size_t begin, size_t end,
                         size_t origin[3], size_t extent[3])
{
  size_t cd[3];
  for (int i = 0; i < 3; i++) {
    if (vol[i] == 0)
      return -1;
    cd[i] = std::min(std::max<size_t>(1, chunk[i]), vol[i]);
  }
  std::vector<Chunk> all;
  try {
    all = chunk_volume(vol, cd);
  }
  catch (const std::exception&) {
    return -1;
  }
  if (begin >= end || end > all.size())
    return -1;
  size_t lo[3] = {~size_t(0), ~size_t(0), ~size_t(0)}, hi[3] = {0, 0, 0};
  for (size_t i = begin; i < end; i++) {
    const Chunk& c = all[i];
    const size_t a[3] = {c.x0, c.y0, c.z0}, l[3] = {c.lx, c.ly, c.lz};
    for (int k = 0; k < 3; k++) {
      lo[k] = std::min(lo[k], a[k]);
      hi[k] = std::max(hi[k], a[k] + l[k]);
    }
  }
  for (int k = 0; k < 3; k++) {
    origin[k] = lo[k];
    extent[k] = hi[k] - lo[k];
  }
  return 0;
}

int sperr_b200_shard_ranges(const size_t vol[3], const size_t chunk[3], size_t world, size_t* begins)
{
  return shard_ranges(vol, chunk, world, begins) ? 0 : -1;
}

int sperr_b200_comp_3d_range_dev(const void* d_box, int is_float, const size_t vol[3],
                                 const size_t chunk[3], const size_t box_origin[3],
                                 const size_t box_extent[3], size_t chunk_begin, size_t chunk_end,
                                 int mode, double quality, const void** d_streams, size_t* streams_len,
                                 uint32_t* lens)
{
  if (quality <= 0.0 || mode < 1 || mode > 3)
    return 2;
  std::lock_guard<std::mutex> lock(g_rmutex);
  return guarded([&] {
    std::vector<Chunk> chunks;
    if (!range_chunks(vol, chunk, chunk_begin, chunk_end, box_origin, box_extent, chunks))
      return -1;
    if (!g_rcomp)
      g_rcomp = &shared_compressor();
    cudaStream_t st = 0;
    SrcVol sv{d_box, is_float, box_extent[0], box_extent[1], box_extent[2]};
    g_rout.reserve(size_t(1) << 20);
    std::vector<size_t> l;
    g_rcomp->max_batch = 0;   // shared with the host-pointer API: no leftovers of its overlap hooks
    g_rcomp->before_batch = nullptr;
    g_rcomp->group_chunks = 0;
    g_rcomp->before_group = nullptr;
    g_rcomp->compress(sv, chunks, mode, quality, false, g_rout, l, st);
    size_t total = 0;
    for (size_t i = 0; i < l.size(); i++) {
      if (l[i] > 0xFFFFFFFFull)
        return -1;
      lens[i] = uint32_t(l[i]);
      total += l[i];
    }
    rt::sync(st);
    *d_streams = g_rout.p;   // library-owned device buffer, valid until the next call
    *streams_len = total;
    return 0;
  });
}

size_t sperr_b200_container_header(const size_t vol[3], const size_t chunk[3], int is_float,
                                   const uint32_t* lens, size_t nchunks, void* out, size_t cap)
{
  const size_t hlen = (nchunks > 1 ? 20 : 14) + 4 * nchunks;
  if (!out || cap < hlen)
    return hlen;
  uint8_t* o = static_cast<uint8_t*>(out);
  o[0] = 0;  // SPERR_VERSION_MAJOR
  o[1] = uint8_t(0x40 | (is_float ? 0x20 : 0) | (nchunks > 1 ? 0x10 : 0));
  const uint32_t v3[3] = {uint32_t(vol[0]), uint32_t(vol[1]), uint32_t(vol[2])};
  std::memcpy(o + 2, v3, 12);
  size_t pos = 14;
  if (nchunks > 1) {
    uint16_t c3[3];
    for (int i = 0; i < 3; i++)
      c3[i] = uint16_t(std::min(std::max<size_t>(1, chunk[i]), vol[i]));
    std::memcpy(o + pos, c3, 6);
    pos += 6;
  }
  std::memcpy(o + pos, lens, 4 * nchunks);
  return hlen;
}

int sperr_b200_parse_container(const void* src, size_t len, size_t vol[3], size_t chunk[3],
                               int* is_float, size_t* header_len, uint32_t* lens, size_t cap,
                               size_t* nchunks)
{
  const uint8_t* p = static_cast<const uint8_t*>(src);
  if (!p || len < 14 || p[0] != 0 || !(p[1] & 0x40))
    return -1;
  const bool multi = (p[1] & 0x10) != 0;
  *is_float = (p[1] & 0x20) ? 1 : 0;
  uint32_t v3[3];
  std::memcpy(v3, p + 2, 12);
  size_t pos = 14;
  for (int i = 0; i < 3; i++)
    vol[i] = chunk[i] = v3[i];
  if (multi) {
    if (len < 20)
      return -1;
    uint16_t c3[3];
    std::memcpy(c3, p + 14, 6);
    for (int i = 0; i < 3; i++)
      chunk[i] = c3[i];
    pos = 20;
  }
  for (int i = 0; i < 3; i++)
    if (vol[i] == 0 || chunk[i] == 0)
      return -1;
  size_t nseg[3], n = 0;
  if (!chunk_grid(vol, chunk, nseg, &n) || n > (~size_t(0) - pos) / 4)
    return -1;
  *nchunks = n;
  *header_len = pos + 4 * n;
  if (len < pos + 4 * n)
    return -1;
  if (lens) {
    if (cap < n)
      return -1;
    std::memcpy(lens, p + pos, 4 * n);
  }
  return 0;
}

int sperr_b200_memcpy_dev(void* dst, const void* src, size_t n, int kind)
{
  return guarded([&] {
    cudaStream_t st = 0;
    if (kind == 0)
      rt::d2d(dst, src, n, st);
    else if (kind == 1)
      rt::h2d(dst, src, n, st);
    else
      rt::d2h(dst, src, n, st);
    rt::sync(st);
    return 0;
  });
}

int sperr_b200_decomp_3d_range_dev(const void* h_streams, const void* d_streams, size_t streams_len,
                                   const uint32_t* lens,
                                   const size_t vol[3], const size_t chunk[3],
                                   const size_t box_origin[3], const size_t box_extent[3],
                                   size_t chunk_begin, size_t chunk_end, int output_float,
                                   void* d_box_out)
{
  std::lock_guard<std::mutex> lock(g_rmutex);
  return guarded([&] {
    std::vector<Chunk> chunks;
    if (!range_chunks(vol, chunk, chunk_begin, chunk_end, box_origin, box_extent, chunks))
      return -1;
    std::vector<ChunkStream> cs(chunks.size());
    size_t off = 0;
    for (size_t i = 0; i < chunks.size(); i++) {
      cs[i].off = off;
      cs[i].len = lens[i];
      off += lens[i];
    }
    if (off != streams_len)
      return -1;
    if (!g_rdecomp)
      g_rdecomp = &shared_decompressor();
    cudaStream_t st = 0;
    const uint8_t* ds = static_cast<const uint8_t*>(d_streams);
    if (!ds && !h_streams)
      return -1;
    if (!ds) {
      g_rstream.reserve(streams_len + 16);
      rt::h2d(g_rstream.p, h_streams, streams_len, st);
      ds = g_rstream.as<uint8_t>();
    }
    SrcVol dv{d_box_out, output_float, box_extent[0], box_extent[1]};
    g_rdecomp->max_batch = 0;
    g_rdecomp->after_batch = nullptr;
    g_rdecomp->group_chunks = 0;
    g_rdecomp->after_group = nullptr;
    g_rdecomp->multires = nullptr;
    g_rdecomp->decompress(static_cast<const uint8_t*>(h_streams), ds, chunks, cs, dv, st);
    rt::sync(st);
    return 0;
  });
}

}  // extern "C"
