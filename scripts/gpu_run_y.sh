#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
timeout 900 python -m pytest tests/test_gpu_fer_curve.py -q -m gpu 2>&1 | tail -15
