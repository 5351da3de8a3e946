#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python bench.py --prec i16 --no-configs --no-cpu --no-int16 --e2e-sockets llr8_packed --steps 5 2>gpurun_out/r2_i16e2e.err > gpurun_out/r2_i16e2e.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_i16e2e.json"))
print(d["value"], json.dumps(d["e2e"])[:1500])
PY
tail -5 gpurun_out/r2_i16e2e.err | cut -c1-300
