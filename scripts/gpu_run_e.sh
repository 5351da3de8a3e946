#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ldpc.py tests/test_gpu_golden.py -q -m gpu -x 2>&1 | tail -5
for v in "--frames 113664" "--workload 8psk89 --frames 113664" "--workload qpsk35 --frames 113664" "--workload 16apsk89 --frames 113664"; do
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e $v 2>> gpurun_out/bench_e.err | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print(d['config']['workload'], d['dtype'], d['config']['frames_per_step_per_gpu'], d['value'], d['stage_ms'], d['roofline']['message_passing'], d['fer']['mean_iterations'], d['fer']['monitor']['fer'])"
done
tail -3 gpurun_out/bench_e.err
