#!/usr/bin/env bash
# round 2: one 12-warp CTA against two 6-warp CTAs per SM for the LDPC kernel — device-resident and through four e2e clones
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
for c in 1 2; do
  export DVBS2_LDPC_CTAS_PER_SM=$c
  timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-sockets sc16_packed,f32_int32 --steps 6 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); e=d['e2e']; print('CTAs/SM $c: device %9.1f ldpc %7.2f | e2e %s  clone0 %s' % (d['value'], d['stage_ms']['ldpc'], {k: v.get('value') for k, v in e['sockets'].items()}, e.get('clone0_stage_ms_per_call')))"
  timeout 300 python bench.py --no-configs --no-cpu --no-int16 --no-e2e --frames 113664 --steps 6 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('CTAs/SM $c: single wave device %9.1f ldpc %7.2f' % (d['value'], d['stage_ms']['ldpc']))"
  timeout 300 python bench.py --prec i16 --no-configs --no-cpu --no-int16 --e2e-sockets llr8_packed --steps 5 2>/dev/null | python -c "
import json,sys; d=json.load(sys.stdin); print('CTAs/SM $c: int16 device %9.1f | e2e %s' % (d['value'], d['e2e']['value']))"
done
unset DVBS2_LDPC_CTAS_PER_SM
timeout 900 python -m pytest -q -m gpu -x tests/test_gpu_ldpc.py tests/test_gpu_parity_at_scale.py -k "ragged or compaction or three_round or four_chunk" 2>&1 | tail -2
