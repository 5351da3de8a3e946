#!/usr/bin/env bash
# round 2, fourth GPU pass: on-chip LDPC kernel v2 (prefetch stages) — parity, timing
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_bch_demod_chain.py -q -x -k "f32_bit_exact or ragged or extreme or normalisation or no_early or demodulate" 2>&1 | tail -5 | tee gpurun_out/r2_tests_d.log
timeout -k 10 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity_at_scale.py tests/test_gpu_fer_curve.py -q -x -k "256 or onchip or monte" 2>&1 | tail -5 | tee -a gpurun_out/r2_tests_d.log
for wl in qpsk89 qpsk35 8psk89; do
  timeout 600 python bench.py --workload $wl --onchip 1 --no-configs --no-cpu --no-int16 --no-e2e --steps 5 2>gpurun_out/r2_onchip2_$wl.err > gpurun_out/r2_onchip2_$wl.json; tail -2 gpurun_out/r2_onchip2_$wl.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_onchip2_*.json")):
    try:
        d = json.load(open(f))
        print(f, d["value"], d["stage_ms"], d["fer"]["mean_iterations"])
    except Exception as ex:
        print(f, "failed", ex)
PY
