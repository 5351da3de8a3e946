#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/exp_t.jsonl
run() { echo "## $*" >> gpurun_out/exp_t.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --no-int16 --steps 5 --warmup 3 "$@" >> gpurun_out/exp_t.jsonl 2>> gpurun_out/exp_t.err; }
run --rounds 4 --exit-eighths 6 --nf-later 2
run --rounds 4 --exit-eighths 7 --nf-later 2
run --rounds 4 --exit-eighths 7
run --rounds 3 --exit-eighths 7 --nf-later 2
run --rounds 3 --exit-eighths 6 --nf-later 2
run --rounds 4 --exit-eighths 5 --nf-later 2
run --frames 227328 --rounds 4 --exit-eighths 6 --nf-later 2
run --frames 113664 --rounds 4 --exit-eighths 6 --nf-later 2
run --frames 113664 --rounds 2 --exit-eighths 6
run --prec i16 --rounds 4 --exit-eighths 6
run --prec i16 --rounds 3 --exit-eighths 6
run --prec i16 --rounds 4 --exit-eighths 7
run --prec i16 --rounds 4 --exit-eighths 6 --nf-later 4
python - <<'PY'
import json
for line in open("gpurun_out/exp_t.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["config"]["frames_per_step_per_gpu"], d["stage_ms"]["ldpc"], round(d["fer"]["mean_iterations"], 3), d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_t.err
