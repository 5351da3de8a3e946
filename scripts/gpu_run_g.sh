#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 7 -c 1 -o gpurun_out/prof_ldpc_round2 -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --frames 113664 > gpurun_out/ncu_full_run.log 2>&1
tail -2 gpurun_out/ncu_full_run.log | cut -c1-200
