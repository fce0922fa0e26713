"""bench.py's host logic (timed loops, stage ranges, JSON line) executed in this GPU-less container.

The device layer is faked for the run: torch's CUDA calls become no-ops / host timers, the process
group is gloo, and the compute behind the C ABI is the CPU SIMT emulation of the CUDA kernels
(tests/emul -- test infrastructure only; the product never loads it). What is checked is that
`python bench.py` walks through every statement of its own arm and prints one JSON line with the
keys the driver reads; numbers are meaningless here.
"""
import json
import os
import subprocess
import sys

import gpulib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

DRIVER = r"""
import sys, time, types
sys.path.insert(0, %(root)r)
sys.path.insert(0, %(tests)r)
import torch
import torch.distributed as dist

class Ev:
    def __init__(self, enable_timing=True):
        self.t = None
    def record(self):
        self.t = time.perf_counter()
    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3

cpu = torch.device("cpu")
real_device = torch.device
torch.cuda.Event = Ev
torch.cuda.set_device = lambda *a, **k: None
torch.cuda.synchronize = lambda *a, **k: None
torch.device = lambda *a, **k: cpu
real_init = dist.init_process_group
dist.init_process_group = lambda backend, rank, world_size, device_id=None: real_init("gloo", rank=rank, world_size=world_size)

import sperr_b200
from sperr_b200 import api
real_load = api.load
sperr_b200.load = lambda path=None: real_load(%(emul)r)

import bench
sys.argv = ["bench.py", "--size", "32", "--steps", "2", "--warmup", "1", "--e2e", "0", "--cpu-baseline", "0", "--settle", "0"]
bench.main()
"""


def test_bench_own_arm_dry_run():
    gpulib.build_emul()
    code = DRIVER % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "emul": gpulib.EMUL_SO}
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29617", RANK="0", WORLD_SIZE="1", LOCAL_RANK="0")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-3000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1]
    d = json.loads(line)
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "clocks", "e2e", "gpu_launches",
              "stages_ms", "compress_gbs", "decompress_gbs", "step_ms_each", "compress_ms_each",
              "decompress_ms_each"):
        assert k in d, k
    assert d["steps"] == 2 and len(d["step_ms_each"]) == 2 and len(d["decompress_ms_each"]) == 2
    assert d["max_abs_err"] <= 1e-3 + 1.2e-7


def test_bench_two_ranks_dry_run():
    """the torchrun shape of the same arm: two ranks, chunk streams gathered to rank 0 (gloo here)"""
    gpulib.build_emul()
    code = DRIVER % {"root": ROOT, "tests": os.path.join(ROOT, "tests"), "emul": gpulib.EMUL_SO}
    code = code.replace('"--size", "32"', '"--gpus", "2", "--size", "32"')
    code = code.replace("bench.main()", "bench.CHUNK = 16   # 32^3 values: 8 chunks, 4 per rank\nbench.main()")
    procs = []
    for rank in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29627", RANK=str(rank), WORLD_SIZE="2",
                   LOCAL_RANK=str(rank))
        procs.append(subprocess.Popen([sys.executable, "-c", code], stdout=subprocess.PIPE, stderr=subprocess.PIPE,
                                      text=True, env=env))
    outs = [p.communicate(timeout=900) for p in procs]
    for p, (so, se) in zip(procs, outs):
        assert p.returncode == 0, se[-3000:]
    lines = [ln for ln in outs[0][0].splitlines() if ln.startswith("{")]
    assert lines and not [ln for ln in outs[1][0].splitlines() if ln.startswith("{")]   # rank 0 alone prints
    d = json.loads(lines[-1])
    assert d["n_gpus"] == 2 and d["scaling"] == "strong" and len(d["step_ms_each"]) == 2
    assert "32x32x32" in d["config"]["workload"]           # the SAME volume at every N
    assert d["parity"]["n_gpus"] == 2 and len(d["parity"]["container_sha256_16"]) == 16
    assert d["extra"]["weak"]["scaling"] == "weak" and "32x32x64" in d["extra"]["weak"]["config"]["workload"]


def test_bench_helpers_without_gpu():
    """the pieces of bench.py that talk to the OS: host counters (deltas go into the bench line) and
    the clock sampler switched off (--diag noclocks): no thread, no NVML, an empty summary"""
    import importlib
    import sys as _sys
    argv = _sys.argv
    _sys.argv = ["bench.py"]
    try:
        bench = importlib.import_module("bench")
    finally:
        _sys.argv = argv
    hc = bench.host_counters()
    assert isinstance(hc, dict) and all(isinstance(v, int) for v in hc.values())
    clk = bench.Clocks(0, idle=True)
    clk.sample()
    clk.mark()
    s = clk.summary()
    assert s["samples"] == 0 and s["sm_mhz"] is None and s["reasons"] == []
