#!/usr/bin/env bash
# scheduling sweep of the LDPC rounds + one line per BASELINE workload / precision (one gpurun call): results in gpurun_out/
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/sweep_b.jsonl
run() { echo "## $*" >> gpurun_out/sweep_b.jsonl; timeout 300 python bench.py --no-e2e --no-cpu --steps 6 --warmup 3 "$@" >> gpurun_out/sweep_b.jsonl 2>> gpurun_out/sweep_b.err; }
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
run --frames 227328 --rounds 2 --nf-later 2
run --frames 340992 --rounds 3 --nf-later 2
run --no-compaction
run --prec i16
run --prec i8
run --workload qpsk35
run --workload qpsk35 --prec i16
run --workload 8psk35
run --workload 8psk35 --prec i16
run --workload 8psk89
run --workload 16apsk89
python - <<'PY'
import json
for line in open("gpurun_out/sweep_b.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["stage_ms"], round(d["fer"]["mean_iterations"], 3), d["fer"]["bch_fer"], d["gpu_launches"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/sweep_b.err
