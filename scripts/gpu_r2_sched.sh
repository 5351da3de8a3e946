#!/usr/bin/env bash
# round 2: scheduling knobs of the multi-round LDPC decoder re-measured on the current code, same box for all lines
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { # workload, tag, flags...
  wl=$1; tag=$2; shift 2
  timeout 300 python bench.py --workload $wl --no-configs --no-cpu --no-int16 --no-e2e --steps 6 "$@" 2>/dev/null > gpurun_out/r2_sched_${wl}_$tag.json
  python - gpurun_out/r2_sched_${wl}_$tag.json "$wl $tag $*" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-40s %9.1f  %7.2f ms  ldpc %7.2f  it %.3f" % (sys.argv[2], d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"], d["fer"]["mean_iterations"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
for wl in qpsk89 8psk89 qpsk35; do
  run $wl default
  run $wl r4e6n2 --rounds 4 --exit-eighths 6 --nf-later 2
  run $wl r4e5n2 --rounds 4 --exit-eighths 5 --nf-later 2
  run $wl r3e6n2 --rounds 3 --exit-eighths 6 --nf-later 2
  run $wl r4e6n1 --rounds 4 --exit-eighths 6
done
run qpsk89 default2
for wl in qpsk89 8psk89 qpsk35; do
  run $wl i16_default --prec i16
  run $wl i16_r4e7 --prec i16 --rounds 4 --exit-eighths 7
  run $wl i16_r5e7 --prec i16 --rounds 5 --exit-eighths 7
  run $wl i16_r4e6 --prec i16 --rounds 4 --exit-eighths 6
done
timeout 600 python -m pytest -q -m gpu -x tests/test_gpu_compact_sockets.py -k "async" 2>&1 | tail -3
