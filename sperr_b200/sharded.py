"""Chunk-sharded compression / decompression across the GPUs of one box (one process per GPU).

SPERR's chunks are independent (/root/reference/src/SPERR3D_OMP_C.cpp:94-130), so the volume is
partitioned into contiguous ranges of ``chunk_volume``'s order (sperr_helper.cpp:542-592) and every
rank runs the CUDA hot path on its own chunks through the C ABI (include/sperr_b200.h, section 2b).
``torch.distributed`` (NCCL over NVLink on GPUs, gloo in the CPU tests) is only plumbing for the one
exchange each way: all-gather of the per-chunk byte counts, which the container header needs
(SPERR3D_OMP_C.cpp:225-230), and gather / scatter of the compressed chunk streams. The result on
rank 0 is the reference's single-stream container, byte-identical to what ``sperr_comp_3d`` writes
for the whole volume.
"""
import ctypes as C

import numpy as np
import torch
import torch.distributed as dist

sz = C.c_size_t
vp = C.c_void_p
sz3 = sz * 3

_libc = C.CDLL(None)
_libc.free.argtypes = [vp]
_libc.free.restype = None


def _bind(cdll):
    if getattr(cdll, "_sperr_sharded_bound", False):
        return cdll
    cdll.sperr_b200_num_chunks.restype = sz
    cdll.sperr_b200_num_chunks.argtypes = [sz3, sz3]
    cdll.sperr_b200_chunk_box.restype = C.c_int
    cdll.sperr_b200_chunk_box.argtypes = [sz3, sz3, sz, sz, sz3, sz3]
    cdll.sperr_b200_comp_3d_range_dev.restype = C.c_int
    cdll.sperr_b200_comp_3d_range_dev.argtypes = [vp, C.c_int, sz3, sz3, sz3, sz3, sz, sz, C.c_int,
                                                  C.c_double, C.POINTER(vp), C.POINTER(sz), vp]
    cdll.sperr_b200_memcpy_dev.restype = C.c_int
    cdll.sperr_b200_memcpy_dev.argtypes = [vp, vp, sz, C.c_int]
    cdll.sperr_b200_container_header.restype = sz
    cdll.sperr_b200_container_header.argtypes = [sz3, sz3, C.c_int, vp, sz, vp, sz]
    cdll.sperr_b200_parse_container.restype = C.c_int
    cdll.sperr_b200_parse_container.argtypes = [vp, sz, sz3, sz3, C.POINTER(C.c_int), C.POINTER(sz),
                                                vp, sz, C.POINTER(sz)]
    cdll.sperr_b200_decomp_3d_range_dev.restype = C.c_int
    cdll.sperr_b200_decomp_3d_range_dev.argtypes = [vp, vp, sz, vp, sz3, sz3, sz3, sz3, sz, sz, C.c_int, vp]
    cdll._sperr_sharded_bound = True
    return cdll


def chunk_range(nchunks, rank, world):
    """Contiguous range of chunk indices owned by `rank` (earlier ranks get the remainder)."""
    base, rem = divmod(nchunks, world)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


class Shard:
    """What one rank owns: chunk range and the bounding box of those chunks inside the volume."""

    def __init__(self, cdll, vol, chunk, rank, world):
        self.cdll = _bind(cdll)
        self.vol, self.chunk = sz3(*vol), sz3(*chunk)
        self.nchunks = int(self.cdll.sperr_b200_num_chunks(self.vol, self.chunk))
        if self.nchunks < world:
            raise ValueError("fewer chunks (%d) than ranks (%d)" % (self.nchunks, world))
        self.begin, self.end = chunk_range(self.nchunks, rank, world)
        self.origin, self.extent = sz3(), sz3()
        rc = self.cdll.sperr_b200_chunk_box(self.vol, self.chunk, self.begin, self.end, self.origin,
                                            self.extent)
        assert rc == 0
        self.ranges = [chunk_range(self.nchunks, r, world) for r in range(world)]

    @property
    def box_origin(self):
        return tuple(self.origin)

    @property
    def box_extent(self):
        return tuple(self.extent)

    def slices(self):
        """numpy / torch index of the box inside a (z, y, x) view of the whole volume."""
        o, e = self.box_origin, self.box_extent
        return (slice(o[2], o[2] + e[2]), slice(o[1], o[1] + e[1]), slice(o[0], o[0] + e[0]))


def _world(group):
    return dist.get_rank(group), dist.get_world_size(group)


_pinned = {}


def _pinned_u8(n, key):
    """Grow-only pinned staging (one buffer per use per process)."""
    t = _pinned.get(key)
    if t is None or t.numel() < n:
        t = torch.empty(max(n, 1 << 20), dtype=torch.uint8)
        if torch.cuda.is_available():
            t = t.pin_memory()
        _pinned[key] = t
    return t


class DeviceContainer:
    """A reference-layout container that stays where it was produced: `data` is a uint8 tensor on
    rank 0's device holding header + chunk streams, `header` the same header bytes on the host
    (they are all a reader needs to find the chunks). `numpy()` brings the whole thing to the host."""

    def __init__(self, data, header):
        self.data, self.header = data, header

    @property
    def size(self):
        return int(self.data.numel())

    def numpy(self):
        return self.data.cpu().numpy()


def compress_3d_sharded(cdll, box, vol, chunk, mode, quality, group=None, device_container=False):
    """box: this rank's part of the volume as a contiguous float32 / float64 torch tensor (z, y, x)
    that lives where the library computes (CUDA device for libsperr_b200.so). Returns the container
    on rank 0 and None elsewhere: a uint8 numpy array backed by a reused pinned staging buffer, or --
    device_container=True -- a DeviceContainer whose bytes never leave rank 0's device."""
    rank, world = _world(group)
    sh = Shard(cdll, vol, chunk, rank, world)
    assert box.is_contiguous() and tuple(box.shape) == sh.box_extent[::-1], (box.shape, sh.box_extent)
    is_float = box.dtype == torch.float32
    n_mine = sh.end - sh.begin
    lens = np.zeros(n_mine, dtype=np.uint32)
    d_streams, n = vp(None), sz(0)
    rc = sh.cdll.sperr_b200_comp_3d_range_dev(vp(box.data_ptr()), int(is_float), sh.vol, sh.chunk,
                                              sh.origin, sh.extent, sh.begin, sh.end, mode, quality,
                                              C.byref(d_streams), C.byref(n), lens.ctypes.data_as(vp))
    dev = box.device
    if world == 1:
        # a world of one rank needs no exchange: header + this rank's streams are the container
        if rc != 0:
            raise RuntimeError("sperr_b200_comp_3d_range_dev failed (rc=%d)" % rc)
        hlen = int(sh.cdll.sperr_b200_container_header(sh.vol, sh.chunk, int(is_float), None, sh.nchunks, None, 0))
        total = hlen + n.value
        hdr = np.zeros(hlen, dtype=np.uint8)
        got = sh.cdll.sperr_b200_container_header(sh.vol, sh.chunk, int(is_float), lens.ctypes.data_as(vp),
                                                  sh.nchunks, hdr.ctypes.data_as(vp), hdr.size)
        assert got == hlen
        if device_container:
            data = torch.empty(total, dtype=torch.uint8, device=dev)
            data[:hlen].copy_(torch.from_numpy(hdr))
            if n.value:
                assert sh.cdll.sperr_b200_memcpy_dev(vp(data.data_ptr() + hlen), d_streams, n.value, 0) == 0
            return DeviceContainer(data, hdr)
        stage = _pinned_u8(total, "container")
        out = stage.numpy()[:total]
        out[:hlen] = hdr
        if n.value:
            kind = 2 if dev.type == "cuda" else 0   # device -> host (the emulated library "device" is the host)
            assert sh.cdll.sperr_b200_memcpy_dev(vp(stage.data_ptr() + hlen), d_streams, n.value, kind) == 0
        return out
    ok = torch.tensor([abs(rc)], device=dev, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MAX, group=group)
    if int(ok.item()) != 0:
        raise RuntimeError("sperr_b200_comp_3d_range_dev failed on some rank (rc=%d here)" % rc)

    # 1. all-gather the per-chunk byte counts (ranges may differ by one chunk: pad)
    per = max(e - b for b, e in sh.ranges)
    mylens = torch.zeros(per, dtype=torch.int64, device=dev)
    mylens[:n_mine] = torch.from_numpy(lens.astype(np.int64)).to(dev)
    all_lens = [torch.zeros(per, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_lens, mylens, group=group)
    lens_by_rank = [t.cpu().numpy()[:e - b] for t, (b, e) in zip(all_lens, sh.ranges)]
    bytes_by_rank = [int(l.sum()) for l in lens_by_rank]

    # 2. gather the chunk streams on rank 0, device to device (variable length: pad to the longest)
    longest = max(max(bytes_by_rank), 1)
    payload = torch.empty(longest, dtype=torch.uint8, device=dev)
    if n.value:
        rc = sh.cdll.sperr_b200_memcpy_dev(vp(payload.data_ptr()), d_streams, n.value, 0)
        assert rc == 0
    parts = [torch.empty(longest, dtype=torch.uint8, device=dev) for _ in range(world)] if rank == 0 else None
    dist.gather(payload, parts, dst=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    if rank != 0:
        return None

    # 3. reference-layout container: header with every chunk length, then the streams in chunk order
    all32 = np.concatenate(lens_by_rank).astype(np.uint32)
    hlen = int(sh.cdll.sperr_b200_container_header(sh.vol, sh.chunk, int(is_float), None, sh.nchunks, None, 0))
    total = hlen + sum(bytes_by_rank)
    if device_container:
        hdr = np.zeros(hlen, dtype=np.uint8)
        got = sh.cdll.sperr_b200_container_header(sh.vol, sh.chunk, int(is_float), all32.ctypes.data_as(vp),
                                                  sh.nchunks, hdr.ctypes.data_as(vp), hdr.size)
        assert got == hlen
        data = torch.empty(total, dtype=torch.uint8, device=dev)
        data[:hlen].copy_(torch.from_numpy(hdr))
        pos = hlen
        for r in range(world):
            data[pos:pos + bytes_by_rank[r]].copy_(parts[r][:bytes_by_rank[r]])
            pos += bytes_by_rank[r]
        return DeviceContainer(data, hdr)
    stage = _pinned_u8(total, "container")
    out = stage.numpy()[:total]
    got = sh.cdll.sperr_b200_container_header(sh.vol, sh.chunk, int(is_float), all32.ctypes.data_as(vp),
                                              sh.nchunks, out.ctypes.data_as(vp), out.size)
    assert got == hlen
    pos = hlen
    for r in range(world):
        stage[pos:pos + bytes_by_rank[r]].copy_(parts[r][:bytes_by_rank[r]])
        pos += bytes_by_rank[r]
    if dev.type == "cuda":
        torch.cuda.current_stream(dev).synchronize()
    return out


def parse_container(cdll, stream):
    """(vol, chunk, is_float, header_len, lens) of a container held in host memory."""
    cdll = _bind(cdll)
    stream = np.ascontiguousarray(stream, dtype=np.uint8)
    vol, chunk = sz3(), sz3()
    isf, hlen, nch = C.c_int(0), sz(0), sz(0)
    rc = cdll.sperr_b200_parse_container(stream.ctypes.data_as(vp), stream.size, vol, chunk, C.byref(isf),
                                         C.byref(hlen), None, 0, C.byref(nch))
    if rc != 0:
        raise ValueError("not a SPERR 3D container")
    lens = np.zeros(nch.value, dtype=np.uint32)
    rc = cdll.sperr_b200_parse_container(stream.ctypes.data_as(vp), stream.size, vol, chunk, C.byref(isf),
                                         C.byref(hlen), lens.ctypes.data_as(vp), lens.size, C.byref(nch))
    if rc != 0 or hlen.value + int(lens.astype(np.int64).sum()) != stream.size:
        raise ValueError("truncated SPERR 3D container")
    return tuple(vol), tuple(chunk), bool(isf.value), hlen.value, lens


def parse_header_only(cdll, header, total_len):
    """parse_container for a container whose header bytes are on the host and whose total length is
    known (the chunk streams may live elsewhere)."""
    cdll = _bind(cdll)
    header = np.ascontiguousarray(header, dtype=np.uint8)
    vol, chunk = sz3(), sz3()
    isf, hlen, nch = C.c_int(0), sz(0), sz(0)
    rc = cdll.sperr_b200_parse_container(header.ctypes.data_as(vp), header.size, vol, chunk, C.byref(isf),
                                         C.byref(hlen), None, 0, C.byref(nch))
    if rc != 0 or hlen.value != header.size:
        raise ValueError("not a SPERR 3D container header")
    lens = np.zeros(nch.value, dtype=np.uint32)
    rc = cdll.sperr_b200_parse_container(header.ctypes.data_as(vp), header.size, vol, chunk, C.byref(isf),
                                         C.byref(hlen), lens.ctypes.data_as(vp), lens.size, C.byref(nch))
    if rc != 0 or hlen.value + int(lens.astype(np.int64).sum()) != total_len:
        raise ValueError("truncated SPERR 3D container")
    return tuple(vol), tuple(chunk), bool(isf.value), hlen.value, lens


def _decompress_world1(cdll, stream, dev, output_float, on_device):
    """decompress_3d_sharded for a world of one rank: no exchange at all."""
    if on_device:
        vol, chunk, isf, hlen, lens = parse_header_only(cdll, stream.header, stream.size)
        d_all = stream.data
    else:
        vol, chunk, isf, hlen, lens = parse_container(cdll, stream)
        d_all = torch.from_numpy(np.ascontiguousarray(stream)).to(dev)
    sh = Shard(cdll, vol, chunk, 0, 1)
    nb = int(lens.astype(np.int64).sum())
    mine = d_all[hlen:hlen + nb]
    mylens = np.ascontiguousarray(lens, dtype=np.uint32)
    # device container: no host copy of the streams, the library fetches the chunk headers itself
    h = None if on_device else np.ascontiguousarray(stream)[hlen:hlen + nb]
    e = sh.box_extent
    box = torch.empty((e[2], e[1], e[0]), dtype=torch.float32 if output_float else torch.float64, device=dev)
    rc = cdll.sperr_b200_decomp_3d_range_dev(h.ctypes.data_as(vp) if h is not None else vp(None),
                                             vp(mine.data_ptr()), nb,
                                             mylens.ctypes.data_as(vp), sh.vol, sh.chunk, sh.origin,
                                             sh.extent, sh.begin, sh.end, int(output_float),
                                             vp(box.data_ptr()))
    if rc != 0:
        raise RuntimeError("sperr_b200_decomp_3d_range_dev failed (rc=%d)" % rc)
    return box, sh


def decompress_3d_sharded(cdll, stream, device, output_float=True, group=None):
    """stream: the container (uint8 numpy array, or the DeviceContainer compress_3d_sharded returned)
    on rank 0, ignored elsewhere. Every rank returns
    (box, shard): its part of the decoded volume as a (z, y, x) tensor on `device`, and the Shard that
    says where the box sits."""
    rank, world = _world(group)
    cdll = _bind(cdll)
    dev = torch.device(device)
    src0 = dist.get_global_rank(group, 0) if group is not None else 0
    on_device = isinstance(stream, DeviceContainer)
    if world == 1:
        return _decompress_world1(cdll, stream, dev, output_float, on_device)
    meta = [None]
    if rank == 0:
        if on_device:
            vol, chunk, isf, hlen, lens = parse_header_only(cdll, stream.header, stream.size)
        else:
            vol, chunk, isf, hlen, lens = parse_container(cdll, stream)
        meta = [(vol, chunk, lens)]
    dist.broadcast_object_list(meta, src=src0, group=group)
    vol, chunk, lens = meta[0]
    sh = Shard(cdll, vol, chunk, rank, world)
    bytes_by_rank = [int(lens[b:e].astype(np.int64).sum()) for b, e in sh.ranges]
    longest = max(max(bytes_by_rank), 1)
    mine = torch.empty(longest, dtype=torch.uint8, device=dev)
    parts = None
    if rank == 0:
        # one upload of the container, then equal-sized device slices for the scatter
        d_all = stream.data if on_device else torch.from_numpy(np.ascontiguousarray(stream)).to(dev)
        parts, pos = [], hlen
        for r in range(world):
            t = torch.empty(longest, dtype=torch.uint8, device=dev)
            t[:bytes_by_rank[r]].copy_(d_all[pos:pos + bytes_by_rank[r]])
            parts.append(t)
            pos += bytes_by_rank[r]
    dist.scatter(mine, parts, src=src0, group=group)
    nb = bytes_by_rank[rank]
    mylens = np.ascontiguousarray(lens[sh.begin:sh.end], dtype=np.uint32)
    meta_flags = [bool(on_device)]
    dist.broadcast_object_list(meta_flags, src=src0, group=group)
    if meta_flags[0]:
        # the streams stay on the device: the library fetches the chunk headers (conditioner 17 B,
        # SPECK header 9 B, outlier header 9 B) it parses on the host itself
        h = None
    else:
        # bring this rank's streams over once (pinned staging)
        stage = _pinned_u8(nb, "streams")
        stage[:nb].copy_(mine[:nb])
        if dev.type == "cuda":
            torch.cuda.current_stream(dev).synchronize()
        h = stage.numpy()[:nb]
    e = sh.box_extent
    box = torch.empty((e[2], e[1], e[0]), dtype=torch.float32 if output_float else torch.float64,
                      device=dev)
    rc = cdll.sperr_b200_decomp_3d_range_dev(h.ctypes.data_as(vp) if h is not None else vp(None),
                                             vp(mine.data_ptr()), nb,
                                             mylens.ctypes.data_as(vp), sh.vol, sh.chunk, sh.origin,
                                             sh.extent, sh.begin, sh.end, int(output_float),
                                             vp(box.data_ptr()))
    ok = torch.tensor([abs(rc)], device=dev, dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MAX, group=group)
    if int(ok.item()) != 0:
        raise RuntimeError("sperr_b200_decomp_3d_range_dev failed (rc=%d here)" % rc)
    return box, sh
