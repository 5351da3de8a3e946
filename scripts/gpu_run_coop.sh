#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -12
: > gpurun_out/exp_coop.jsonl
run() { echo "## $*" >> gpurun_out/exp_coop.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --steps 8 --warmup 3 "$@" >> gpurun_out/exp_coop.jsonl 2>> gpurun_out/exp_coop.err; }
run --frames 1024 --e2e-clones 1
run --frames 4096 --e2e-clones 1
run --frames 16384 --e2e-clones 1
run --frames 32768 --e2e-clones 8
run --frames 32768 --e2e-clones 1
run --frames 65536 --e2e-clones 4
python - <<'PY'
import json
for line in open("gpurun_out/exp_coop.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            e = d["e2e"]
            print("dev", d["value"], d["ms_per_step"], "| e2e", e["value"], e["ms_per_step"], e["clones"], "x", e["frames_per_step_per_gpu"] // e["clones"], "packed", e["packed_output"]["value"], d["fer"]["mean_iterations"])
        except Exception as ex:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_coop.err
