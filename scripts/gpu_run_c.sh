#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_c.json 2> gpurun_out/bench_c.err; tail -5 gpurun_out/bench_c.err; cat gpurun_out/bench_c.json
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --workload 8psk35 >> gpurun_out/bench_c_variants.json 2>> gpurun_out/bench_c.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-clones 2 >> gpurun_out/bench_c_variants.json 2>> gpurun_out/bench_c.err
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu --e2e-clones 8 >> gpurun_out/bench_c_variants.json 2>> gpurun_out/bench_c.err
cat gpurun_out/bench_c_variants.json
timeout 1500 python -m pytest tests -q -m gpu -k "tx or demod" 2>&1 | tail -5
