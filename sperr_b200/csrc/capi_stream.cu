// Volumes that do not fit beside the coder's work buffers: sperr_comp_3d / sperr_decomp_3d take them
// through HBM one group of chunk slabs at a time, the way the reference walks its chunk list with a
// few chunks in flight (/root/reference/src/SPERR3D_OMP_C.cpp:84-141, SPERR3D_OMP_D.cpp:84-126).
// Only the current group's part of the volume is on the device; the container is assembled on the
// host from the groups' streams. Chunks are numbered z-major (chunk_volume), so a group of whole
// z-slabs of chunks is both a contiguous chunk range and a contiguous part of the host volume, and
// the chunk-range entry points (capi_range.cu) do the work.
//
// Taken when the volume is larger than SPERR_B200_STREAM_MB (a test aid: any size can be forced
// through this path) or than 45 % of the device's memory; the group size is a quarter of the device
// memory (or the given limit). Returns -2 when the call cannot be split that way (one slab of chunks
// is already too large, or the volume is not a whole number of slabs): the caller then takes the
// resident path, which fails with -1 when memory runs out, as before.
#include "../../include/sperr_b200.h"

#include <cstring>
#include <mutex>

#include "hostpipe.h"
#include "pipeline.h"

using namespace sperr_b200;

namespace sperr_b200 {

namespace {

struct StreamState {   // per device
  rt::DBuf box;
  std::mutex mu;   // one streamed call per device at a time (the chunk-range calls inside take the API mutex)
};
StreamState& stream_state()
{
  static StreamState s[rt::kMaxDevices];
  return s[rt::cur_dev()];
}

// bytes of the volume one group may hold on the device; 0: the resident path is fine
size_t stream_group_limit(size_t volume_bytes)
{
  if (const char* e = std::getenv("SPERR_B200_STREAM_MB")) {
    const size_t lim = size_t(std::max(1, std::atoi(e))) << 20;
    return volume_bytes > lim ? lim : 0;
  }
#ifndef SPERR_EMUL
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  if (volume_bytes > total_b / 100 * 45)
    return total_b / 4;
#endif
  return 0;
}

// groups of whole z-slabs of chunks: [first chunk, last chunk) and [z0, z1) of every group
struct Group {
  size_t c0, c1, z0, z1;
};
bool slab_groups(const size_t vol[3], const size_t cd[3], size_t esz, size_t limit, std::vector<Group>& out)
{
  const auto chunks = chunk_volume(vol, cd);
  if (chunks.empty())
    return false;
  size_t per_slab = 0;
  while (per_slab < chunks.size() && chunks[per_slab].z0 == chunks[0].z0)
    per_slab++;
  const size_t nslabs = chunks.size() / per_slab;
  if (per_slab * nslabs != chunks.size() || nslabs < 2)
    return false;
  const size_t plane = vol[0] * vol[1] * esz;
  size_t s0 = 0;
  while (s0 < nslabs) {
    size_t s1 = s0 + 1;
    auto z_of = [&](size_t s) { return s == nslabs ? vol[2] : size_t(chunks[s * per_slab].z0); };
    if ((z_of(s1) - z_of(s0)) * plane > limit)
      return false;   // a single slab of chunks does not fit
    while (s1 < nslabs && (z_of(s1 + 1) - z_of(s0)) * plane <= limit)
      s1++;
    out.push_back(Group{s0 * per_slab, s1 * per_slab, z_of(s0), z_of(s1)});
    s0 = s1;
  }
  return out.size() >= 2;
}

}  // namespace

int comp_3d_streamed(const void* src, int is_float, const size_t vol[3], const size_t chunk[3], int mode,
                     double quality, void** dst, size_t* dst_len)
{
  const size_t esz = is_float ? 4 : 8;
  const size_t limit = stream_group_limit(vol[0] * vol[1] * vol[2] * esz);
  if (!limit)
    return -2;
  size_t cd[3];
  for (int i = 0; i < 3; i++)
    cd[i] = std::min(std::max<size_t>(1, chunk[i]), vol[i]);
  std::vector<Group> groups;
  try {
    if (!slab_groups(vol, cd, esz, limit, groups))
      return -2;
  }
  catch (const std::exception&) {
    return -1;
  }
  const size_t nchunks = groups.back().c1;
  const size_t plane = vol[0] * vol[1] * esz;
  std::vector<uint32_t> lens(nchunks, 0);
  std::vector<std::vector<uint8_t>> streams(groups.size());
  try {
    StreamState& s = stream_state();
    std::lock_guard<std::mutex> lock(s.mu);
    for (size_t g = 0; g < groups.size(); g++) {
      const Group& gr = groups[g];
      const size_t org[3] = {0, 0, gr.z0}, ext[3] = {vol[0], vol[1], gr.z1 - gr.z0};
      const size_t bytes = (gr.z1 - gr.z0) * plane;
      s.box.reserve(bytes);
      HostPipe::get().h2d(s.box.p, static_cast<const char*>(src) + gr.z0 * plane, bytes, 0);
      const void* d_streams = nullptr;
      size_t n = 0;
      const int rc = sperr_b200_comp_3d_range_dev(s.box.p, is_float, vol, cd, org, ext, gr.c0, gr.c1, mode, quality,
                                                  &d_streams, &n, lens.data() + gr.c0);
      if (rc != 0)
        return rc;
      streams[g].resize(n);
      if (n) {
        rt::d2h(streams[g].data(), d_streams, n, 0);
        rt::sync(0);
      }
    }
  }
  catch (const std::exception& e) {
    if (std::getenv("SPERR_B200_VERBOSE"))
      std::fprintf(stderr, "sperr_b200 (streamed): %s\n", e.what());
    return -1;
  }
  const size_t hlen = sperr_b200_container_header(vol, cd, is_float, nullptr, nchunks, nullptr, 0);
  size_t total = hlen;
  for (const auto& v : streams)
    total += v.size();
  uint8_t* const container = static_cast<uint8_t*>(std::malloc(total));
  if (!container)
    return -1;
  sperr_b200_container_header(vol, cd, is_float, lens.data(), nchunks, container, hlen);
  size_t off = hlen;
  for (const auto& v : streams) {
    if (!v.empty())
      std::memcpy(container + off, v.data(), v.size());
    off += v.size();
  }
  *dst = container;
  *dst_len = total;
  return 0;
}

int decomp_3d_streamed(const void* src, size_t src_len, int output_float, size_t* dimx, size_t* dimy,
                       size_t* dimz, void** dst)
{
  size_t vol[3], cd[3], hlen = 0, nchunks = 0;
  int is_float = 0;
  if (sperr_b200_parse_container(src, src_len, vol, cd, &is_float, &hlen, nullptr, 0, &nchunks) != 0)
    return -1;
  const size_t esz = output_float ? 4 : 8;
  if (vol[0] == 0 || vol[1] == 0 || vol[2] == 0)
    return -2;
  const size_t limit = stream_group_limit(vol[0] * vol[1] * vol[2] * esz);
  if (!limit)
    return -2;
  std::vector<uint32_t> lens(nchunks);
  if (sperr_b200_parse_container(src, src_len, vol, cd, &is_float, &hlen, lens.data(), nchunks, &nchunks) != 0)
    return -1;
  std::vector<size_t> coff(nchunks + 1, hlen);
  for (size_t i = 0; i < nchunks; i++)
    coff[i + 1] = coff[i] + lens[i];
  if (coff[nchunks] != src_len)
    return -1;
  std::vector<Group> groups;
  try {
    if (!slab_groups(vol, cd, esz, limit, groups) || groups.back().c1 != nchunks)
      return -2;
  }
  catch (const std::exception&) {
    return -1;
  }
  const size_t plane = vol[0] * vol[1] * esz;
  void* const out = std::malloc(vol[0] * vol[1] * vol[2] * esz);
  if (!out)
    return -1;
  try {
    StreamState& s = stream_state();
    std::lock_guard<std::mutex> lock(s.mu);
    for (const Group& gr : groups) {
      const size_t org[3] = {0, 0, gr.z0}, ext[3] = {vol[0], vol[1], gr.z1 - gr.z0};
      const size_t bytes = (gr.z1 - gr.z0) * plane;
      s.box.reserve(bytes);
      const size_t b0 = coff[gr.c0], b1 = coff[gr.c1];
      const int rc = sperr_b200_decomp_3d_range_dev(static_cast<const uint8_t*>(src) + b0, nullptr, b1 - b0,
                                                    lens.data() + gr.c0, vol, cd, org, ext, gr.c0, gr.c1, output_float,
                                                    s.box.p);
      if (rc != 0) {
        std::free(out);
        return rc;
      }
      HostPipe::get().d2h(static_cast<char*>(out) + gr.z0 * plane, s.box.p, bytes, 0);
      HostPipe::get().wait_idle();
    }
  }
  catch (const std::exception& e) {
    if (std::getenv("SPERR_B200_VERBOSE"))
      std::fprintf(stderr, "sperr_b200 (streamed): %s\n", e.what());
    std::free(out);
    return -1;
  }
  *dimx = vol[0];
  *dimy = vol[1];
  *dimz = vol[2];
  *dst = out;
  return 0;
}

}  // namespace sperr_b200
