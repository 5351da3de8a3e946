#!/usr/bin/env bash
# 2-GPU: C++ module check + golden tests on GPU 0, then the bench exactly as the driver launches it at N=2 (both arms)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m pytest tests/test_gpu_cpp_modules.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 2> gpurun_out/bench_mg2.err | tee gpurun_out/bench_mg2.json | cut -c1-2500
tail -3 gpurun_out/bench_mg2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-400
