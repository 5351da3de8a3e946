#!/usr/bin/env bash
# e2e sweep: module clones x frames per call
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/exp_p.jsonl
run() { echo "## $*" >> gpurun_out/exp_p.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --steps 6 --warmup 3 --frames 113664 "$@" >> gpurun_out/exp_p.jsonl 2>> gpurun_out/exp_p.err; }
run --e2e-clones 2
run --e2e-clones 3
run --e2e-clones 4
run --e2e-clones 6
run --e2e-clones 3 --e2e-frames 56832
run --e2e-clones 4 --e2e-frames 227328
run --e2e-clones 2 --e2e-frames 227328
python - <<'PY'
import json
for line in open("gpurun_out/exp_p.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            e = d["e2e"]
            print(e["value"], e["ms_per_step"], e["frames_per_step_per_gpu"], "packed", e["packed_output"]["value"])
        except Exception as ex:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_p.err
