#!/usr/bin/env bash
# round 2: the cooperative small-batch LDPC kernel for int16 — parity (every int16 test now runs both kernels), then call times
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -m gpu -x tests/test_gpu_ldpc.py tests/test_gpu_golden.py tests/test_gpu_compact_sockets.py tests/test_gpu_bch_demod_chain.py tests/test_gpu_cpp_modules.py 2>&1 | tail -3
for F in 1024 4096 16384; do
  for cm in 0 -1; do
    timeout 300 python bench.py --prec i16 --frames $F --coop-max $cm --no-configs --no-cpu --no-int16 --no-e2e --steps 10 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('int16 %6d frames coop-max %3s: %8.3f ms per call  %9.1f Mbit/s  ldpc %7.3f ms  launches/step %d' % ($F, '$cm', d['ms_per_step'], d['value'], d['stage_ms']['ldpc'], d['gpu_launches'] // d['steps']))"
  done
done
