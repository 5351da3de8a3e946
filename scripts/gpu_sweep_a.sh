#!/usr/bin/env bash
# GPU tests + scheduling sweep of the LDPC rounds (one gpurun call): results in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/sweep_a_gpu.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/sweep_a_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/sweep_a_tests.log
tail -5 gpurun_out/sweep_a_tests.log
: > gpurun_out/sweep_a.jsonl
run() { echo "## $*" >> gpurun_out/sweep_a.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --steps 6 --warmup 3 "$@" >> gpurun_out/sweep_a.jsonl 2>> gpurun_out/sweep_a.err; }
run
run --rounds 3
run --rounds 3 --nf-later 2
run --rounds 2 --nf-later 2
run --rounds 4
run --rounds 3 --exit-eighths 5
run --rounds 3 --exit-eighths 3
run --frames 227328
run --frames 227328 --rounds 3
run --frames 227328 --rounds 3 --nf-later 2
run --frames 170496 --rounds 3
run --no-compaction
python - <<'PY'
import json
for line in open("gpurun_out/sweep_a.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["stage_ms"], d["fer"]["mean_iterations"], d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/sweep_a.err
