#!/usr/bin/env bash
# multi-GPU bench exactly as the driver launches it; stdout of the bench must be ONE JSON line
cd "${GRAFT_REPO_ROOT:-.}"
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${2:-4} --warmup 3 ${3:-} 2> gpurun_out/bench_mg$N.err > gpurun_out/bench_mg$N.json
echo "stdout lines: $(grep -c . gpurun_out/bench_mg$N.json)"; cut -c1-300 gpurun_out/bench_mg$N.json
tail -3 gpurun_out/bench_mg$N.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-200
