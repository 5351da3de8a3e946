#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -5
timeout 900 python bench.py 2> gpurun_out/bench_r.err | tee gpurun_out/bench_r.json | cut -c1-3800
tail -3 gpurun_out/bench_r.err
timeout 300 python bench.py --prec i16 --no-e2e --no-cpu --steps 5 2>&1 | tail -1 | cut -c1-1200
