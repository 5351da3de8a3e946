#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
./scripts/membench_bulk > gpurun_out/membench_bulk.txt 2>&1; cat gpurun_out/membench_bulk.txt
timeout 900 python -m pytest tests/test_gpu_tx_monitor.py -x -q -m gpu 2>&1 | tail -5
