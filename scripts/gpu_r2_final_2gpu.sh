#!/usr/bin/env bash
# round 2: the final bench exactly as the driver launches it on 2 GPUs (both arms)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
N=2
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r2_final_bench_mg$N.err > gpurun_out/r2_final_bench_mg$N.json
echo "stdout lines: $(grep -c . gpurun_out/r2_final_bench_mg$N.json)"
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final_bench_mg2.json"))
e = d["e2e"]
print(d["value"], d["n_gpus"], "e2e", e["value"], e["headline_sockets"], {k: (v.get("value"), v.get("clones"), v.get("host_gbs_total"), v.get("error")) for k, v in e["sockets"].items()}, d["wall_s"])
PY
tail -3 gpurun_out/r2_final_bench_mg$N.err | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300
