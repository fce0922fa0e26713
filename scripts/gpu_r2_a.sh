#!/bin/bash
# round 2, call A (2 GPUs): NCCL world-2 byte identity, strong-scaled bench at N=1 and N=2, the
# reference arm under torchrun
mkdir -p gpurun_out
export SPERR_B200_VERBOSE=1
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_gpu_sharded.py -x -q > gpurun_out/r2a_pytest.log 2>&1; tail -3 gpurun_out/r2a_pytest.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/r2a_bench1.log 2>&1; tail -1 gpurun_out/r2a_bench1.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2a_bench2.log 2>&1; tail -1 gpurun_out/r2a_bench2.log | cut -c1-1500
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2a_ref2.log 2>&1; tail -1 gpurun_out/r2a_ref2.log | cut -c1-600
