#!/usr/bin/env bash
# small-batch behaviour of the cooperative kernel vs the lane = frame kernel (ldpc_coop_max_frames) on one box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,temperature.gpu,power.draw --format=csv,noheader
: > gpurun_out/exp_coop2.jsonl
run() { echo "## $*" >> gpurun_out/exp_coop2.jsonl; timeout 300 python bench.py --no-cpu --no-int16 --no-e2e --steps 8 --warmup 3 "$@" >> gpurun_out/exp_coop2.jsonl 2>> gpurun_out/exp_coop2.err; }
run --frames 1024
run --frames 4096
run --frames 16384
run --frames 32768
run --frames 1024 --coop-max -1
run --frames 4096 --coop-max -1
run --frames 16384 --coop-max -1
run --frames 32768 --coop-max -1
run --frames 49152 --coop-max 65536
run --frames 49152 --coop-max -1
run --frames 4096 --workload qpsk35
run --frames 4096 --workload 8psk89
python - <<'PY'
import json
for line in open("gpurun_out/exp_coop2.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print("dev", d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"])
        except Exception as ex:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_coop2.err
