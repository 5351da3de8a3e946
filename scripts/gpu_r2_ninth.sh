#!/usr/bin/env bash
# round 2, ninth GPU pass (gpurun --gpus 8): the bench at N = 8 exactly as the driver launches it (device-resident + e2e with the
# compact and the reference's socket formats, NUMA binding reported per rank), and the clone round-robin test on 8 GPUs
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8; nproc; lscpu | grep -E "NUMA node|Socket"; free -g | head -2
nvidia-smi topo -m 2>/dev/null | head -12 | cut -c1-160
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-configs --e2e-sockets sc16_packed,f32_int32 2> gpurun_out/r2_bench_mg$N.err > gpurun_out/r2_bench_mg$N.json
echo "stdout lines: $(grep -c . gpurun_out/r2_bench_mg$N.json)"
python - $N <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_bench_mg%s.json" % sys.argv[1]))
print(d["value"], d["n_gpus"], {k: (v.get("value"), v.get("host_gbs_total"), v.get("error")) for k, v in d["e2e"]["sockets"].items()}, d["e2e"].get("host_numa"), d.get("wall_s"))
PY
tail -3 gpurun_out/r2_bench_mg$N.err | cut -c1-200
timeout 300 python -m pytest -q -m gpu -x tests/test_gpu_cpp_modules.py -k "round_robin" 2>&1 | tail -1
