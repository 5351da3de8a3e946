#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
free -g | head -2
timeout 1500 python bench.py 2> gpurun_out/r2_final2_bench.err > gpurun_out/r2_final2_bench.json
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2_final2_bench.json"))
print(d["value"], d["wall_s"], "e2e", d["e2e"]["value"], d["e2e"]["vs_device_resident"], {k: v.get("value") for k, v in d["e2e"]["sockets"].items()})
print("int16", d["int16"]["value"], json.dumps(d["int16"].get("e2e"))[:1200])
print({k: v.get("value") for k, v in d["configs"].items()})
PY
tail -3 gpurun_out/r2_final2_bench.err | cut -c1-300
