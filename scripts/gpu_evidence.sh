#!/usr/bin/env bash
# round-end evidence: full GPU suite, smoke, default bench (both arms), ncu launch lists (f32, int16)
cd "${GRAFT_REPO_ROOT:-.}"
T=${1:-final}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_$T.err | tee gpurun_out/bench_$T.json | cut -c1-3800
tail -3 gpurun_out/bench_$T.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_${T}_ref.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_f32.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-int16 > gpurun_out/launches_$T.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${T}_i16.csv \
  python bench.py --prec i16 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_${T}_i16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 6 -c 1 -o gpurun_out/prof_ldpc_${T}_i16 -f \
  python bench.py --prec i16 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_${T}_i16.log 2>&1
ls -la gpurun_out | tail -8
