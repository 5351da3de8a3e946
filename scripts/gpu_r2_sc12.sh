#!/usr/bin/env bash
# round 2: 12-bit packed sample sockets (sc12) — parity with the float sockets on the same sample values, then the e2e legs
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -x tests/test_gpu_compact_sockets.py tests/test_gpu_rx_front.py 2>&1 | tail -2
timeout 900 python bench.py --no-configs --no-cpu --no-int16 --e2e-sockets sc16_packed,sc12_packed,sc8_packed --steps 6 2>gpurun_out/r2_sc12.err > gpurun_out/r2_sc12.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_sc12.json"))
e = d["e2e"]
print("device", d["value"], "| e2e", e.get("value"), e.get("vs_device_resident"), e.get("headline_sockets"), d["wall_s"])
for k, v in e["sockets"].items():
    print("  ", k, v.get("value"), v.get("clones"), v.get("frames_per_step_per_gpu"), v.get("ms_per_step"), v.get("host_gbs_per_gpu"), v.get("bch_fer"), v.get("clone0_stage_ms_per_call"), v.get("error"))
PY
tail -3 gpurun_out/r2_sc12.err | cut -c1-300
