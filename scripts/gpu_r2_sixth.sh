#!/usr/bin/env bash
# round 2, sixth GPU pass: racecheck with indirect barrier dependencies (the mbarrier protocols have waiters that never arrive),
# the rebuilt reference binaries with the GPU stage (tests + --sim-stats tables at three -F)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck --racecheck-indirect-barrier-dependency yes --log-file gpurun_out/r2_racecheck_indirect.log --print-limit 20 \
  python -m pytest -q -m gpu -x "tests/test_gpu_ldpc.py::test_ragged_batches[33]" 2>&1 | tail -2
tail -4 gpurun_out/r2_racecheck_indirect.log
timeout 900 python -m pytest -q -m gpu -x tests/test_gpu_binaries.py tests/test_gpu_cpp_modules.py tests/test_host_side.py -s 2>&1 | grep -E "passed|failed|dvbs2_|Error" | tail -8
R=baseline/_ref
if [ -x $R/bin/gpu/dvbs2_tx_rx_bb ]; then
  mkdir -p $R/run; cd $R/run
  export LD_LIBRARY_PATH=$GRAFT_REPO_ROOT/dvbs2_b200:$LD_LIBRARY_PATH
  for cfg in "2048 4" "4096 8" "16384 8" "16384 16"; do
    set -- $cfg
    timeout 600 ../bin/gpu/dvbs2_tx_rx_bb --mod-cod QPSK-S_8/9 --dec-implem NMS --dec-ite 10 -m 4.0 -M 4.05 --ter-freq 0 --dec-hw GPU \
      --gpu-sockets FLOAT -F $1 --sim-threads $2 -e 600 --sim-stats > $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_F$1_t$2.txt 2>&1
    grep -E "demodulate_decode|^ +[0-9.]+ +\|" $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_F$1_t$2.txt | tail -3
  done
  timeout 600 ../bin/cpu/dvbs2_tx_rx_bb --mod-cod QPSK-S_8/9 --dec-implem NMS --dec-ite 10 -m 4.0 -M 4.05 --ter-freq 0 --dec-simd INTER -F 16 -e 300 --sim-stats > $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_cpu.txt 2>&1
  grep -E "decode_siho|decode_hiho|demodulate|^ +[0-9.]+ +\|" $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_cpu.txt | tail -6
  cd $GRAFT_REPO_ROOT
fi
