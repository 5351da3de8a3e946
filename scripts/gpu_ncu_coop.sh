#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coop -s 2 -c 1 -o gpurun_out/prof_ldpc_coop -f \
  python bench.py --frames 4096 --steps 2 --warmup 2 --no-e2e --no-cpu --no-int16 > gpurun_out/ncu_coop.log 2>&1
tail -2 gpurun_out/ncu_coop.log | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_coop4096.csv \
  python bench.py --frames 4096 --steps 2 --warmup 1 --no-e2e --no-cpu --no-int16 > gpurun_out/launches_coop.log 2>&1
ls -la gpurun_out | tail -4
