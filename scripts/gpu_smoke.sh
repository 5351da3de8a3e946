cd "${GRAFT_REPO_ROOT:-.}"
timeout 120 python __graft_entry__.py smoke 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_bch_demod_chain.py -q -m gpu -x -k "fused_chain or empty" 2>&1 | tail -1
