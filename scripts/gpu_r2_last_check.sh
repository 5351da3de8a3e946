#!/usr/bin/env bash
# round 2: last sanity pass after the async guard — compact sockets (incl. the async entry point), the C++ modules, the binaries, smoke
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_compact_sockets.py tests/test_gpu_cpp_modules.py tests/test_gpu_bch_demod_chain.py tests/test_gpu_binaries.py 2>&1 | tail -2
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
