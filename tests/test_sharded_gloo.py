"""Host-side logic of the chunk-sharded path (sperr_b200/sharded.py) with world_size 2 over gloo:
partitioning, all-gather of the chunk lengths, gather / scatter of the chunk streams and container
assembly. The compute behind the C ABI is the CPU SIMT emulation of the CUDA kernels (tests/emul,
test infrastructure only) because this container has no GPU; on the GPU box the same module runs on
libsperr_b200.so over NCCL (bench.py --gpus N, tests/test_gpu_sharded.py)."""
import ctypes as C
import os
import random
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import gpulib
import refs

CASES = [
    # vol, chunk, mode, quality
    ((48, 40, 36), (16, 16, 16), 3, 1e-3),     # 3*2*2 = 12 chunks, some ragged
    ((64, 32, 32), (32, 32, 32), 2, 70.0),     # exactly one chunk per rank
    ((40, 24, 20), (16, 16, 16), 1, 3.0),      # odd count: 2*1*1 ... ranges differ in size
]


def _worker(rank, world, port, case, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from sperr_b200 import sharded
        vol, chunk, mode, quality = case
        cdll = C.CDLL(gpulib.EMUL_SO)
        v = refs.synthetic_field(vol, seed=11).reshape(vol[2], vol[1], vol[0])
        sh = sharded.Shard(cdll, vol, chunk, rank, world)
        box = torch.from_numpy(np.ascontiguousarray(v[sh.slices()]))
        stream = sharded.compress_3d_sharded(cdll, box, vol, chunk, mode, quality)
        oracle = refs.oracle()
        if rank == 0:
            rc, exp = oracle.comp_3d(v.reshape(-1), vol, chunk, mode, quality)
            assert rc == 0
            assert stream.size == exp.size and np.array_equal(stream, exp), "container differs"
        else:
            assert stream is None
        out, sh2 = sharded.decompress_3d_sharded(cdll, stream, torch.device("cpu"), True)
        rc, exp = oracle.comp_3d(v.reshape(-1), vol, chunk, mode, quality)
        rc, dec, dims = oracle.decomp_3d(exp, True)
        assert rc == 0 and dims == tuple(vol)
        want = dec.reshape(vol[2], vol[1], vol[0])[sh2.slices()]
        assert np.array_equal(out.numpy().view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
        # the same with the container kept where it was produced: only the chunk headers visit the host
        dc = sharded.compress_3d_sharded(cdll, box, vol, chunk, mode, quality, device_container=True)
        if rank == 0:
            assert isinstance(dc, sharded.DeviceContainer) and np.array_equal(dc.numpy(), exp), "device container differs"
        else:
            assert dc is None
        out3, sh3 = sharded.decompress_3d_sharded(cdll, dc, torch.device("cpu"), True)
        assert np.array_equal(out3.numpy().view(np.uint32), np.ascontiguousarray(want).view(np.uint32))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, "FAIL: %r" % (e,)))
    finally:
        dist.destroy_process_group()


def test_sharded_world3_on_2x2x2_grid_gloo():
    """8 chunks over 3 ranks: an even split of the chunk order (3, 3, 2) is not a set of boxes, so
    the partition falls back to whole rows (2, 1, 1 rows): every rank's box holds only its own
    chunks and the assembled container / decoded boxes are still the oracle's."""
    gpulib.load("emul")
    refs.oracle()
    from sperr_b200 import sharded
    cdll = C.CDLL(gpulib.EMUL_SO)
    seen = []
    for r in range(3):
        sh = sharded.Shard(cdll, (32, 32, 32), (16, 16, 16), r, 3)
        e = sh.box_extent
        assert e[0] * e[1] * e[2] == (sh.end - sh.begin) * 16 ** 3, "box holds foreign chunks"
        seen.extend(range(sh.begin, sh.end))
    assert seen == list(range(8))
    with pytest.raises(ValueError):
        sharded.Shard(cdll, (16, 48, 32), (16, 16, 16), 0, 4)   # 1 x 3 x 2 grid, 4 ranks: no box split
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = random.randint(20000, 40000)
    case = ((32, 32, 32), (16, 16, 16), 3, 1e-3)
    procs = [ctx.Process(target=_worker, args=(r, 3, port, case, q)) for r in range(3)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


@pytest.mark.parametrize("case", CASES, ids=lambda c: "x".join(map(str, c[0])) + "-m%d" % c[2])
def test_sharded_roundtrip_gloo_world2(case):
    gpulib.load("emul")   # builds tests/emul/libsperr_emul.so when needed
    refs.oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = random.randint(20000, 40000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res


def test_sharded_world1_gloo():
    """a world of one rank takes the exchange-free path (what bench.py --gpus 1 runs)"""
    gpulib.build_emul()
    refs.oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    p = ctx.Process(target=_worker, args=(0, 1, random.randint(20000, 40000), CASES[0], q))
    p.start()
    res = q.get(timeout=300)
    p.join(timeout=60)
    assert res[1] == "ok", res


def test_chunk_ranges_cover_everything():
    from sperr_b200 import sharded
    for n in (1, 2, 7, 8, 64, 65):
        for w in (1, 2, 3, 8):
            got = []
            for r in range(w):
                b, e = sharded.chunk_range(n, r, w)
                got.extend(range(b, e))
            assert got == list(range(n))
