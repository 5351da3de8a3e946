#!/usr/bin/env bash
# round 2, tenth GPU pass: e2e call shapes with ONE chunk per call (multi-wave launches per clone), sc16 sockets
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for cfg in "5 113664" "6 113664" "8 113664" "6 85248" "8 56832"; do
  set -- $cfg
  timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-clones $1 --e2e-frames $2 --e2e-chunks 1 --e2e-sockets sc16_packed --steps 6 2>gpurun_out/r2_e2e1_$1_$2.err > gpurun_out/r2_e2e1_$1_$2.json || tail -3 gpurun_out/r2_e2e1_$1_$2.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_e2e1_*.json")):
    try:
        d = json.load(open(f))
        e = d["e2e"]
        print(f, d["value"], d["config"]["frames_per_step_per_gpu"], "e2e", e["value"], e["ms_per_step"], e["frames_per_step_per_gpu"], e["vs_device_resident"], d.get("wall_s"))
    except Exception as ex:
        print(f, "failed", ex)
PY
