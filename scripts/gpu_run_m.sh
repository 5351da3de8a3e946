#!/usr/bin/env bash
# round-1 final evidence: full GPU suite, smoke, default bench (both arms), ncu launch list, ncu --set full of the f32 and
# packed-int16 LDPC kernels, dram bytes of every LDPC launch of one step
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -4
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py 2> gpurun_out/bench_m.err | tee gpurun_out/bench_m.json | cut -c1-3500
tail -3 gpurun_out/bench_m.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_m_ref.json | cut -c1-400
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_m.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --no-int16 > gpurun_out/launches_m.log 2>&1
grep -c . gpurun_out/launches_m.csv
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_m_i16.csv \
  python bench.py --prec i16 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/launches_m_i16.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 6 -c 1 -o gpurun_out/prof_ldpc_m_f32 -f \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-int16 > gpurun_out/ncu_full_m.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode -s 6 -c 1 -o gpurun_out/prof_ldpc_m_i16 -f \
  python bench.py --prec i16 --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/ncu_full_m_i16.log 2>&1
ls -la gpurun_out | tail -12
