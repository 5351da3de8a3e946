#!/usr/bin/env bash
# int16: ALU-lean pass 2 + narrower later rounds — parity, then sweep
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_golden.py tests/test_gpu_bch_demod_chain.py -q -m gpu -x 2>&1 | tail -12
: > gpurun_out/exp_o.jsonl
run() { echo "## $*" >> gpurun_out/exp_o.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 "$@" >> gpurun_out/exp_o.jsonl 2>> gpurun_out/exp_o.err; }
run --prec i16
run --prec i16 --nf-later 4
run --prec i16 --rounds 2
run --prec i16 --rounds 4
run --prec i16 --frames 227328
run --prec i16 --frames 227328 --rounds 3
run --prec i16 --frames 682240
run --prec i16 --workload qpsk35
run --prec i16 --workload 8psk35
run --prec i16 --workload 8psk89
python - <<'PY'
import json
for line in open("gpurun_out/exp_o.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["config"]["frames_per_step_per_gpu"], d["stage_ms"], round(d["fer"]["mean_iterations"], 3), d["fer"]["bch_fer"], d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_o.err
