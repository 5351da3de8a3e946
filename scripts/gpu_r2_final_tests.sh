#!/usr/bin/env bash
# round 2: the whole GPU suite + smoke on the final code (after the bench-only and sc12 passes)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -q -m gpu --durations=5 2>&1 | tail -10 | tee gpurun_out/r2_final_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
