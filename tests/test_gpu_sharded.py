"""GPU: the chunk-sharded API (sperr_b200/sharded.py) over NCCL with a world of one rank -- the same
code path bench.py --gpus N takes -- produces the oracle's container and decodes it bit-exactly."""
import os
import random

import numpy as np
import pytest

import refs

pytestmark = pytest.mark.gpu


def test_sharded_nccl_world1(oracle):
    import torch
    import torch.distributed as dist

    import sperr_b200
    from sperr_b200 import sharded

    L = sperr_b200.load()
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(random.randint(20000, 40000))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
    try:
        vol, chunk = (128, 96, 80), (64, 64, 64)
        v = refs.synthetic_field(vol, seed=3)
        box = torch.from_numpy(v.reshape(vol[2], vol[1], vol[0])).to(dev)
        for mode, q in ((3, 1e-3), (1, 2.0)):
            stream = sharded.compress_3d_sharded(L.lib, box, vol, chunk, mode, q)
            rc, exp = oracle.comp_3d(v, vol, chunk, mode, q)
            assert rc == 0 and np.array_equal(stream, exp)
            out, sh = sharded.decompress_3d_sharded(L.lib, stream, dev, True)
            rc, dec, dims = oracle.decomp_3d(exp, True)
            assert sh.box_extent == vol
            assert np.array_equal(out.cpu().numpy().reshape(-1).view(np.uint32), dec.view(np.uint32))
            # container kept in HBM: same bytes, same decode, only the chunk headers visit the host
            dc = sharded.compress_3d_sharded(L.lib, box, vol, chunk, mode, q, device_container=True)
            assert dc.data.is_cuda and np.array_equal(dc.numpy(), exp)
            out2, _ = sharded.decompress_3d_sharded(L.lib, dc, dev, True)
            assert torch.equal(out.view(torch.int32), out2.view(torch.int32))
    finally:
        dist.destroy_process_group()


def _nccl_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    import sperr_b200
    from sperr_b200 import sharded

    try:
        os.environ["MASTER_ADDR"] = "127.0.0.1"
        os.environ["MASTER_PORT"] = str(port)
        torch.cuda.set_device(rank)
        dev = torch.device("cuda", rank)
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
        L = sperr_b200.load()
        oracle = refs.oracle()
        # ragged volume: 2 x 2 x 3 chunk grid (12 chunks, 6 per rank), and a 2 x 2 x 2 grid
        for vol, chunk, mode, q_ in (((128, 96, 150), (64, 64, 48), 3, 1e-3), ((64, 64, 64), (32, 32, 32), 1, 3.0)):
            v = refs.synthetic_field(vol, seed=5).reshape(vol[2], vol[1], vol[0])
            sh = sharded.Shard(L.lib, vol, chunk, rank, world)
            box = torch.from_numpy(np.ascontiguousarray(v[sh.slices()])).to(dev)
            rc, exp = oracle.comp_3d(v.reshape(-1), vol, chunk, mode, q_)
            rc2, dec, dims = oracle.decomp_3d(exp, True)
            assert rc == 0 and rc2 == 0
            want = np.ascontiguousarray(dec.reshape(vol[2], vol[1], vol[0])[sh.slices()])
            for device_container in (False, True):
                s = sharded.compress_3d_sharded(L.lib, box, vol, chunk, mode, q_, device_container=device_container)
                if rank == 0:
                    got = s.numpy() if device_container else s
                    assert np.array_equal(got, exp), "container assembled over NCCL differs from the oracle's"
                else:
                    assert s is None
                out, sh2 = sharded.decompress_3d_sharded(L.lib, s, dev, True)
                assert np.array_equal(out.cpu().numpy().view(np.uint32), want.view(np.uint32))
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, "FAIL: %r\n%s" % (e, traceback.format_exc())))
    finally:
        try:
            dist.destroy_process_group()
        except Exception:
            pass


def test_sharded_nccl_world2_bytes_equal_oracle():
    """Two ranks on two GPUs over NCCL (skipped on a one-GPU box; `gpurun --gpus 2`): the container
    rank 0 assembles from the gathered chunk streams is the oracle's container byte for byte, and
    every rank decodes its box bit-exactly."""
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    refs.oracle()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = random.randint(20000, 40000)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
