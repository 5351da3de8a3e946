#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests/test_gpu_bch_demod_chain.py tests/test_gpu_golden.py tests/test_gpu_tx_monitor.py -q -m gpu -x 2>&1 | tail -4
: > gpurun_out/exp_bch.jsonl
run() { echo "## $*" >> gpurun_out/exp_bch.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --no-e2e --steps 8 --warmup 3 "$@" >> gpurun_out/exp_bch.jsonl 2>> gpurun_out/exp_bch.err; }
run --frames 1024
run --frames 4096
run --frames 16384
run
python - <<'PY'
import json
for line in open("gpurun_out/exp_bch.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print("dev", d["value"], d["ms_per_step"], d["stage_ms"])
        except Exception as ex:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_bch.err
