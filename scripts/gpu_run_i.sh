#!/usr/bin/env bash
# chain tests (packed output) + LDPC tests + default bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_bch_demod_chain.py tests/test_gpu_ldpc.py -q -m gpu -x 2>&1 | tail -3
timeout 600 python bench.py 2> gpurun_out/bench_i.err | tee gpurun_out/bench_i.json | cut -c1-3000
tail -3 gpurun_out/bench_i.err
