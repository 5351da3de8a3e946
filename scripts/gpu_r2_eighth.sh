#!/usr/bin/env bash
# round 2, eighth GPU pass: shared-exponential demapper for 8PSK / 16APSK (parity, then the 8PSK / 16APSK bench lines), int16 line after the pass-2 change
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest -q -m gpu -x tests/test_gpu_bch_demod_chain.py tests/test_gpu_golden.py tests/test_gpu_compact_sockets.py tests/test_gpu_rx_front.py tests/test_gpu_cpp_modules.py tests/test_gpu_fer_curve.py 2>&1 | tail -3
for wl in 8psk89 8psk35 16apsk89; do
  timeout 600 python bench.py --workload $wl --no-configs --no-cpu --no-int16 --no-e2e --steps 5 2>gpurun_out/r2_demap2_$wl.err > gpurun_out/r2_demap2_$wl.json; tail -1 gpurun_out/r2_demap2_$wl.err | cut -c1-200
done
timeout 600 python bench.py --prec i16 --no-configs --no-cpu --no-e2e --steps 5 2>gpurun_out/r2_i16_v3.err > gpurun_out/r2_i16_v3.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_demap2_*.json")) + ["gpurun_out/r2_i16_v3.json"]:
    try:
        d = json.load(open(f))
        print(f, d["value"], d["stage_ms"], d["fer"]["mean_iterations"], d["fer"]["bch_fer"])
    except Exception as ex:
        print(f, "failed", ex)
PY
