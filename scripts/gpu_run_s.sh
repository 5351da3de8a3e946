#!/usr/bin/env bash
# f32 scheduling sweep at the default 3-wave batch
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/exp_s.jsonl
run() { echo "## $*" >> gpurun_out/exp_s.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --no-int16 --steps 5 --warmup 3 "$@" >> gpurun_out/exp_s.jsonl 2>> gpurun_out/exp_s.err; }
run
run --rounds 4
run --rounds 3 --exit-eighths 5
run --rounds 4 --exit-eighths 5
run --rounds 3 --exit-eighths 6
run --rounds 4 --exit-eighths 6
run --rounds 4 --exit-eighths 6 --nf-later 2
run --rounds 2
run --workload qpsk35
run --workload qpsk35 --rounds 2
run --workload 8psk89
run --workload 8psk89 --rounds 4 --exit-eighths 6
python - <<'PY'
import json
for line in open("gpurun_out/exp_s.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["config"]["frames_per_step_per_gpu"], d["stage_ms"], round(d["fer"]["mean_iterations"], 3), d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_s.err
