#!/usr/bin/env bash
# run the whole GPU parity suite on the box (use under gpurun)
cd "${GRAFT_REPO_ROOT:-.}"
nvidia-smi --query-gpu=name,memory.total --format=csv
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -30
