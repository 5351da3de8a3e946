#!/usr/bin/env bash
# round-end verification: full GPU suite, smoke, default bench (both arms), e2e with the cooperative kernel up to 32768 frames
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout -k 10 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_v10.err | tee gpurun_out/bench_v10.json | cut -c1-400
tail -3 gpurun_out/bench_v10.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_v10_ref.json | cut -c1-200
timeout 600 python bench.py --no-cpu --no-int16 --coop-max 32768 --steps 6 2> /dev/null | tee gpurun_out/bench_v10_coop32k.json | cut -c1-200
python - <<'PY'
import json
for f in ("bench_v10.json", "bench_v10_coop32k.json"):
    d = json.load(open("gpurun_out/" + f))
    print(f, d["value"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "packed", d["e2e"]["packed_output"]["value"], "int16", (d.get("int16") or {}).get("value"))
PY
