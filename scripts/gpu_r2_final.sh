#!/usr/bin/env bash
# round 2, final evidence pass: whole GPU suite, smoke, the default bench (both arms), ncu launch lists (time + DRAM bytes per
# launch) for the three LDPC shapes the bench reports, ncu --set full of the float LDPC kernel on QPSK-S_8/9 and QPSK-S_3/5,
# memcheck over a slice of the suite that covers every kernel family incl. the round-2 additions
cd "${GRAFT_REPO_ROOT:-.}"
T=${1:-r2_final}
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem --format=csv,noheader
timeout 2400 python -m pytest tests -q -m gpu --durations=8 2>&1 | tail -14 | tee gpurun_out/${T}_tests.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 1500 python bench.py 2> gpurun_out/${T}_bench.err > gpurun_out/${T}_bench.json; cut -c1-1500 gpurun_out/${T}_bench.json; tail -2 gpurun_out/${T}_bench.err | cut -c1-300
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | tee gpurun_out/${T}_bench_ref.json | cut -c1-400
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
B="--steps 2 --warmup 1 --no-e2e --no-cpu --no-int16 --no-configs"
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_qpsk89_f32.csv python bench.py $B > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_qpsk89_i16.csv python bench.py --prec i16 $B > /dev/null 2>&1
timeout 600 ncu --metrics $M --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_qpsk35_f32.csv python bench.py --workload qpsk35 $B > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode_kernel -s 3 -c 1 -o gpurun_out/${T}_prof_ldpc_qpsk89_f32 -f python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu --no-int16 --no-configs > gpurun_out/${T}_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_decode_kernel -s 3 -c 1 -o gpurun_out/${T}_prof_ldpc_qpsk35_f32 -f python bench.py --workload qpsk35 --steps 1 --warmup 1 --no-e2e --no-cpu --no-int16 --no-configs > gpurun_out/${T}_ncu_b.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --log-file gpurun_out/${T}_memcheck.log --print-limit 20 \
  python -m pytest -q -m gpu -x \
  "tests/test_gpu_ldpc.py::test_ragged_batches" \
  "tests/test_gpu_ldpc.py::test_frame_compaction_is_transparent" \
  "tests/test_gpu_ldpc.py::test_multi_round_compaction_fixed_point" \
  "tests/test_gpu_bch_demod_chain.py::test_packed_output_format" \
  "tests/test_gpu_bch_demod_chain.py::test_bch_bit_exact" \
  "tests/test_gpu_bch_demod_chain.py::test_demodulate_and_deinterleave" \
  "tests/test_gpu_compact_sockets.py" \
  "tests/test_gpu_rx_front.py::test_fused_receiver_stage_from_pl_frames" \
  "tests/test_gpu_tx_monitor.py" 2>&1 | tail -3
tail -4 gpurun_out/${T}_memcheck.log
ls -la gpurun_out | grep $T
