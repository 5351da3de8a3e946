#!/usr/bin/env bash
# int16 compaction / scheduling sweep
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ldpc.py -q -m gpu -x -k "compaction or fixed or wrap" 2>&1 | tail -3
: > gpurun_out/exp_k.jsonl
run() { echo "## $*" >> gpurun_out/exp_k.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --steps 5 --warmup 3 "$@" >> gpurun_out/exp_k.jsonl 2>> gpurun_out/exp_k.err; }
run --prec i16
run --prec i16 --frames 454656 --rounds 2
run --prec i16 --frames 454656 --rounds 4
run --prec i16 --frames 454656 --rounds 3 --exit-eighths 5
run --prec i16 --frames 454656 --rounds 3 --exit-eighths 6
run --prec i16 --frames 454656 --rounds 4 --exit-eighths 6
run --prec i16 --frames 227328 --rounds 3
run --prec i16 --frames 227328 --rounds 4 --exit-eighths 6
run --prec i16 --frames 682240 --rounds 3
python - <<'PY'
import json
for line in open("gpurun_out/exp_k.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["config"]["frames_per_step_per_gpu"], d["stage_ms"], round(d["fer"]["mean_iterations"], 3), d["fer"]["bch_fer"], d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_k.err
