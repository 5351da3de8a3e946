#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
./dvbs2_b200/module/_build/module_check "QPSK-S_8/9" 32 3.95
timeout 900 python -m pytest tests/test_gpu_cpp_modules.py -q -m gpu -x 2>&1 | tail -12
