#!/usr/bin/env bash
# round 2: the adaptive multi-round schedule — parity at scale, then adaptive (library default) vs the fixed round-1 default, same box
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest -q -m gpu -x tests/test_gpu_parity_at_scale.py tests/test_gpu_ldpc.py tests/test_gpu_compact_sockets.py -k "scale or compaction or async or ragged" 2>&1 | tail -3
run() { # workload, tag, flags...
  wl=$1; tag=$2; shift 2
  timeout 300 python bench.py --workload $wl --no-configs --no-cpu --no-int16 --no-e2e --steps 6 "$@" 2>/dev/null > gpurun_out/r2_adapt_${wl}_$tag.json
  python - gpurun_out/r2_adapt_${wl}_$tag.json "$wl $tag $*" <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[1]))
    print("%-50s %9.1f  %7.2f ms  ldpc %7.2f  it %.3f  launches %d" % (sys.argv[2], d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"], d["fer"]["mean_iterations"], d["gpu_launches"]))
except Exception as e:
    print(sys.argv[2], "failed", e)
PY
}
FIX="--rounds 3 --exit-eighths 4 --nf-later 1"
for wl in qpsk89 16apsk89 8psk89 qpsk35 8psk35; do
  run $wl adaptive
  run $wl fixed $FIX
done
run qpsk89 i16_adaptive --prec i16
run qpsk89 i16_fixed --prec i16 --rounds 3 --exit-eighths 4 --nf-later 2
run 16apsk89 i16_adaptive --prec i16
run 16apsk89 i16_fixed --prec i16 --rounds 3 --exit-eighths 4 --nf-later 2
