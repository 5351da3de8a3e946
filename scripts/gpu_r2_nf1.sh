#!/usr/bin/env bash
# round 2: 32-frame against 64-frame groups in round 1 at one, two and three waves (device-resident), and what it does to e2e
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
run() { tag=$1; shift
  timeout 300 python bench.py --no-configs --no-cpu --no-int16 --no-e2e --steps 6 "$@" 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('%-46s %9.1f %8.2f ms ldpc %7.2f launches %d' % ('$tag $*', d['value'], d['ms_per_step'], d['stage_ms']['ldpc'], d['gpu_launches']))"
}
for F in 56832 113664 227328 340992; do
  run nf2 --frames $F --nf-first 2
  run nf1 --frames $F --nf-first 1
done
run nf1_r3 --frames 113664 --nf-first 1 --rounds 3
run nf1_r2 --frames 227328 --nf-first 1 --rounds 2
