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
