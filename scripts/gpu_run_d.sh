#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for v in "--e2e-clones 4" "--e2e-clones 3" "--e2e-clones 6" "--e2e-clones 4 --frames 151552"; do
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu $v 2>> gpurun_out/bench_d.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['value'], d['e2e'])"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'dvbs2' -c 60 --csv --log-file gpurun_out/launches_r1.csv \
  python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
grep -c dvbs2 gpurun_out/launches_r1.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 3 -c 1 -o gpurun_out/prof_ldpc_v2 -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
tail -3 gpurun_out/ncu_full_run.log | cut -c1-300
