#!/usr/bin/env bash
# round 2: 12-byte check state for the rate-3/5 codes + the narrowed adaptive rule — full LDPC / chain parity, then the 3/5 lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 2000 python -m pytest -q -m gpu -x tests/test_gpu_ldpc.py tests/test_gpu_golden.py tests/test_gpu_parity_at_scale.py tests/test_gpu_bch_demod_chain.py tests/test_gpu_fer_curve.py tests/test_gpu_cpp_modules.py 2>&1 | tail -3
run() { # workload, tag, flags...
  wl=$1; tag=$2; shift 2
  timeout 300 python bench.py --workload $wl --no-configs --no-cpu --no-int16 --no-e2e --steps 6 "$@" 2>/dev/null > gpurun_out/r2_state3_${wl}_$tag.json
  python - gpurun_out/r2_state3_${wl}_$tag.json "$wl $tag $*" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-40s %9.1f  %7.2f ms  ldpc %7.2f  it %.3f  frames %d  msg-passing %.0f GB/s" % (sys.argv[2], d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"], d["fer"]["mean_iterations"], d["config"]["frames_per_step_per_gpu"], d["roofline"]["message_passing"]["achieved"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
run qpsk35 f32 --frames 340992
run qpsk35 i16 --prec i16 --frames 227328
run 8psk35 f32 --frames 340992
run 8psk35 i16 --prec i16 --frames 454656
run 16apsk89 i16 --prec i16
run qpsk89 f32
