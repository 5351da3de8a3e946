#!/usr/bin/env bash
# round 2: e2e with TWO clones of three-wave calls (each LDPC launch at the efficiency of the device-resident batch, the copies
# of one clone under the kernels of the other) against four clones of single-wave calls, sc16 sockets
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for cfg in "4 113664" "2 340992" "2 227328" "3 227328"; do
  set -- $cfg
  timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-clones $1 --e2e-frames $2 --e2e-chunks 1 --e2e-sockets sc16_packed --steps 6 2>gpurun_out/r2_e2ebig_$1_$2.err > gpurun_out/r2_e2ebig_$1_$2.json
  python - $1 $2 <<'PY'
import json, sys
d = json.load(open("gpurun_out/r2_e2ebig_%s_%s.json" % (sys.argv[1], sys.argv[2])))
e = d["e2e"]
print(sys.argv[1], "x", sys.argv[2], "device", d["value"], "| e2e", e.get("value"), e.get("ms_per_step"), e.get("vs_device_resident"), e.get("clone0_stage_ms_per_call"), json.dumps(e.get("sockets", {}).get("sc16_packed", {}).get("error"))[:300])
PY
done
