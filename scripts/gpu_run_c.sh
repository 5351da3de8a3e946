#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -3 gpurun_out/bench_c.err; cat gpurun_out/bench_c.json
for v in "--prec i16" "--workload qpsk35" "--frames 151552" "--prec i8"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu $v >> gpurun_out/bench_c_variants.json 2>> gpurun_out/bench_c.err
done
cat gpurun_out/bench_c_variants.json
