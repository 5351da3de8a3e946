#!/usr/bin/env bash
# full validation pass: GPU parity suite, smoke, default bench (both arms), ncu launch list, ncu --set full of the LDPC kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
nproc; free -g | sed -n 2p
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 600 python bench.py 2> gpurun_out/bench_h.err | tee gpurun_out/bench_h.json | cut -c1-2500
tail -3 gpurun_out/bench_h.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_h_ref.json | cut -c1-700
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_h.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_h.log 2>&1
grep -c . gpurun_out/launches_h.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 6 -c 2 -o gpurun_out/prof_ldpc_h -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_h.log 2>&1
tail -2 gpurun_out/ncu_full_h.log | cut -c1-200
ls -la gpurun_out
