#!/usr/bin/env bash
# round 2, third GPU pass: the on-chip LDPC kernel (design B) — parity, then timing on the BASELINE workloads; e2e sweep
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout -k 10 900 python -m pytest tests/test_gpu_ldpc.py -q -x -k "f32_bit_exact or ragged or extreme or normalisation or no_early" 2>&1 | tail -15 | tee gpurun_out/r2_tests_c.log
timeout -k 10 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_parity_at_scale.py -q -x -k "256 or onchip" 2>&1 | tail -15 | tee -a gpurun_out/r2_tests_c.log
for wl in qpsk89 qpsk35 8psk89; do
  timeout 600 python bench.py --workload $wl --onchip 1 --no-configs --no-cpu --no-int16 --no-e2e --steps 5 2>gpurun_out/r2_onchip_$wl.err > gpurun_out/r2_onchip_$wl.json; tail -2 gpurun_out/r2_onchip_$wl.err
done
# e2e sweep (sc16 sockets): clones x frames per call x chunks per call
for cfg in "3 113664 2" "3 170496 2" "2 227328 2" "2 227328 4" "3 227328 3" "4 85248 1"; do
  set -- $cfg
  timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-clones $1 --e2e-frames $2 --e2e-chunks $3 --e2e-sockets sc16_packed --steps 6 2>/dev/null > gpurun_out/r2_e2e_$1_$2_$3.json
done
timeout 600 python bench.py --prec i16 --no-configs --no-cpu --no-int16 --e2e-sockets llr8_packed --steps 5 2>gpurun_out/r2_i16_e2e.err > gpurun_out/r2_i16_e2e.json; tail -3 gpurun_out/r2_i16_e2e.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_onchip_*.json")) + sorted(glob.glob("gpurun_out/r2_e2e_*.json")) + ["gpurun_out/r2_i16_e2e.json"]:
    try:
        d = json.load(open(f))
        e = d.get("e2e") or {}
        print(f, d["value"], d["stage_ms"], d["fer"]["mean_iterations"], "e2e", {k: (v.get("value"), v.get("ms_per_step"), v.get("error")) for k, v in (e.get("sockets") or {}).items()})
    except Exception as ex:
        print(f, "failed", ex)
PY
