#!/usr/bin/env bash
# round 2: the e2e call shapes bench.py now picks (float sockets: 4 clones x one wave; compact sockets: 2 clones x the whole batch)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python bench.py --no-configs --no-cpu --no-int16 --steps 6 2>gpurun_out/r2_e2ebig2.err > gpurun_out/r2_e2ebig2.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_e2ebig2.json"))
e = d["e2e"]
print("device", d["value"], "| e2e", e.get("value"), e.get("vs_device_resident"), e.get("api"), d["wall_s"])
for k, v in e["sockets"].items():
    print("  ", k, v.get("value"), v.get("clones"), v.get("frames_per_step_per_gpu"), v.get("ms_per_step"), v.get("host_gbs_per_gpu"), v.get("bch_fer"), v.get("matches_device_path"), v.get("clone0_stage_ms_per_call"), v.get("error"))
PY
tail -3 gpurun_out/r2_e2ebig2.err | cut -c1-300
timeout 600 python bench.py --prec i16 --no-configs --no-cpu --no-int16 --steps 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('int16 device', d['value'], '| e2e', e.get('value'), e.get('vs_device_resident'), e.get('clones'), e.get('frames_per_step_per_gpu'), json.dumps(e.get('sockets', {}).get('llr8_packed', {}).get('error')))"
