#!/usr/bin/env bash
# round 2, seventh GPU pass (gpurun --gpus 2): in-process multi-GPU (StreamPU clones round-robin over the GPUs), fixed-point
# parity after the int16 pass-2 change, racecheck hazard types, and the bench at N = 2 exactly as the driver launches it
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,name --format=csv,noheader; nproc; lscpu | grep -E "NUMA node|Socket"
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_cpp_modules.py -k "round_robin" 2>&1 | tail -2
timeout 900 python -m pytest -q -m gpu -x tests/test_gpu_ldpc.py tests/test_gpu_golden.py tests/test_gpu_compact_sockets.py -k "fixed_point or int16 or normalisation or precisions or 256 or int8" 2>&1 | tail -2
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --log-file gpurun_out/r2_racecheck_all.log --print-limit 12 \
  python -m pytest -q -m gpu -x "tests/test_gpu_ldpc.py::test_ragged_batches[33]" 2>&1 | tail -1
grep -E "Error|Warning|hazard" gpurun_out/r2_racecheck_all.log | cut -c1-220 | head -14
N=2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-configs --e2e-sockets sc16_packed,f32_int32 2> gpurun_out/r2_bench_mg$N.err > gpurun_out/r2_bench_mg$N.json
echo "stdout lines: $(grep -c . gpurun_out/r2_bench_mg$N.json)"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_bench_mg2.json"))
print(d["value"], d["n_gpus"], {k: (v.get("value"), v.get("host_gbs_total")) for k, v in d["e2e"]["sockets"].items()}, d["e2e"].get("host_numa"), d["int16"]["value"] if d.get("int16") else None)
PY
tail -3 gpurun_out/r2_bench_mg$N.err | cut -c1-200
