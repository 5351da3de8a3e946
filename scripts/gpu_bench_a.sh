#!/usr/bin/env bash
# round-1 measurement pass A (run under gpurun): bench line, variants, ncu launch list + full capture of the LDPC kernel
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)" 
set -x
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -3 gpurun_out/bench_a.err; cat gpurun_out/bench_a.json
for v in "--prec i16" "--workload qpsk35" "--frames 37888" "--frames 151552" "--workload 8psk35 --prec i16"; do
  timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu $v >> gpurun_out/bench_a_variants.json 2>> gpurun_out/bench_a.err
done
cat gpurun_out/bench_a_variants.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches_a.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_launch_run.log 2>&1
tail -25 gpurun_out/launches_a.csv
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 3 -c 1 -o gpurun_out/prof_ldpc_a -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_run.log 2>&1
tail -5 gpurun_out/ncu_full_run.log
ls -la gpurun_out
