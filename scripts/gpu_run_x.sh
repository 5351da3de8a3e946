#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_golden.py tests/test_gpu_bch_demod_chain.py -q -m gpu -x 2>&1 | tail -5
: > gpurun_out/exp_x.jsonl
run() { echo "## $*" >> gpurun_out/exp_x.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --no-e2e --steps 6 --warmup 3 "$@" >> gpurun_out/exp_x.jsonl 2>> gpurun_out/exp_x.err; }
run
run --rounds 4 --exit-eighths 6
run --frames 1024
run --frames 4096
run --frames 16384
run --frames 56832
run --frames 113664
run --workload qpsk35
run --workload 8psk89
python - <<'PY'
import json
for line in open("gpurun_out/exp_x.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["config"]["frames_per_step_per_gpu"], d["stage_ms"]["ldpc"], round(d["fer"]["mean_iterations"], 3), d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_x.err
