#!/usr/bin/env bash
# multi-GPU bench exactly as the driver launches it
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 4 --warmup 3 2> gpurun_out/bench_mg$N.err | tee gpurun_out/bench_mg$N.json | cut -c1-1500
tail -5 gpurun_out/bench_mg$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>&1 | tail -2 | cut -c1-600
