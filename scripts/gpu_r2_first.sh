#!/usr/bin/env bash
# round 2, first GPU pass: the whole GPU suite (incl. the at-scale parity tests), smoke, the default bench + its launch list
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv,noheader
nproc; free -g | head -2; lscpu | grep -E "Model name|NUMA node\(s\)|Socket" 
timeout -k 10 2400 python -m pytest tests -q -m gpu -x --durations=12 2>&1 | tail -40 | tee gpurun_out/r2_tests_a.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python bench.py 2> gpurun_out/r2_bench_a.err | tee gpurun_out/r2_bench_a.json | cut -c1-600
tail -5 gpurun_out/r2_bench_a.err
