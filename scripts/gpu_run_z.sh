#!/usr/bin/env bash
# kernel-variant experiment (libraries built under /tmp, selected with DVBS2_CUDA_LIB): padded group stride of the var array
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
: > gpurun_out/exp_z.jsonl
timeout 300 python -m pytest tests/test_gpu_bch_demod_chain.py -q -m gpu -x -k "empty" 2>&1 | tail -2
DVBS2_CUDA_LIB=$PWD/dvbs2_b200/libvar_pad1.so timeout 600 python -m pytest tests/test_gpu_ldpc.py -q -m gpu -x -k "ragged or all_modcods or compaction_is_transparent" 2>&1 | tail -2
for v in "" pad1 pad7 w16 "" pad1 w16; do
  if [ -n "$v" ]; then export DVBS2_CUDA_LIB=$PWD/dvbs2_b200/libvar_$v.so; else unset DVBS2_CUDA_LIB; fi
  echo "## variant=${v:-baseline}" >> gpurun_out/exp_z.jsonl
  timeout 300 python bench.py --no-e2e --no-cpu --no-int16 --steps 6 --warmup 3 >> gpurun_out/exp_z.jsonl 2>> gpurun_out/exp_z.err
done
python - <<'PY'
import json
for line in open("gpurun_out/exp_z.jsonl"):
    if line.startswith("##"):
        print(line.strip(), end="  ->  ")
    else:
        try:
            d = json.loads(line)
            print(d["value"], d["ms_per_step"], d["stage_ms"]["ldpc"])
        except Exception as e:
            print("bad line", line[:200])
PY
tail -5 gpurun_out/exp_z.err
