#!/usr/bin/env bash
# compute-sanitizer memcheck over a small slice of the parity suite (every kernel family once)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/memcheck.log --print-limit 20 \
  python -m pytest -q -m gpu -x \
  "tests/test_gpu_ldpc.py::test_ragged_batches" \
  "tests/test_gpu_ldpc.py::test_frame_compaction_is_transparent" \
  "tests/test_gpu_ldpc.py::test_multi_round_compaction_fixed_point" \
  "tests/test_gpu_bch_demod_chain.py::test_packed_output_format" \
  "tests/test_gpu_bch_demod_chain.py::test_bch_bit_exact" \
  "tests/test_gpu_rx_front.py::test_fused_receiver_stage_from_pl_frames" \
  "tests/test_gpu_tx_monitor.py" 2>&1 | tail -5
tail -15 gpurun_out/memcheck.log
