#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for nc in 4 1; do
timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-sockets sc16_packed --e2e-clones $nc --e2e-frames 113664 --steps 6 2>gpurun_out/r2_e2ediag_$nc.err > gpurun_out/r2_e2ediag_$nc.json
python - $nc <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_e2ediag_%s.json" % sys.argv[1]))
e = d["e2e"]
print("clones", sys.argv[1], "device", d["value"], d["stage_ms"], "| e2e", e["value"], e["ms_per_step"], e.get("clone0_stage_ms_per_call"))
PY
done
timeout 300 python bench.py --no-configs --no-cpu --no-int16 --no-e2e --frames 113664 --steps 6 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('device-resident single wave', d['value'], d['ms_per_step'], d['stage_ms'])"
