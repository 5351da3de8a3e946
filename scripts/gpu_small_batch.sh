#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout -k 10 400 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -4
: > gpurun_out/exp_coop3.jsonl
run() { echo "## $*" >> gpurun_out/exp_coop3.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --no-e2e --steps 8 --warmup 3 "$@" >> gpurun_out/exp_coop3.jsonl 2>> gpurun_out/exp_coop3.err; }
run --frames 1024
run --frames 4096
run --frames 16384
run --frames 32768 --coop-max 65536
run --frames 4096 --workload qpsk35
run --frames 4096 --workload 8psk89
python - <<'PY'
import json
for line in open("gpurun_out/exp_coop3.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print("dev", d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"])
        except Exception as ex:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_coop3.err
