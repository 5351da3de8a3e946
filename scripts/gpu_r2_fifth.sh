#!/usr/bin/env bash
# round 2, fifth GPU pass: racecheck / synccheck of the mbarrier kernels, the patched binary's --sim-stats table,
# ncu --set full of the on-chip LDPC kernel (why design B loses)
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for tool in racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2_$tool.log --print-limit 20 \
    python -m pytest -q -m gpu -x "tests/test_gpu_ldpc.py::test_ragged_batches[33]" 2>&1 | tail -2
  tail -4 gpurun_out/r2_$tool.log
done
R=baseline/_ref
if [ -x $R/bin/gpu/dvbs2_tx_rx_bb ]; then
  mkdir -p $R/run; cd $R/run
  export LD_LIBRARY_PATH=$GRAFT_REPO_ROOT/dvbs2_b200:$LD_LIBRARY_PATH
  for cfg in "2048 4" "4096 8" "16384 8"; do
    set -- $cfg
    timeout 600 ../bin/gpu/dvbs2_tx_rx_bb --mod-cod QPSK-S_8/9 --dec-implem NMS --dec-ite 10 -m 4.0 -M 4.05 --ter-freq 0 --dec-hw GPU \
      --gpu-sockets FLOAT -F $1 --sim-threads $2 -e 600 --sim-stats > $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_F$1_t$2.txt 2>&1
    grep -E "demodulate_decode|Monitor|^ +[0-9.]+ +\|" $GRAFT_REPO_ROOT/gpurun_out/r2_bb_stats_F$1_t$2.txt | tail -4
  done
  cd $GRAFT_REPO_ROOT
fi
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ldpc_onchip -c 1 -o gpurun_out/prof_onchip_v2 -f \
  python bench.py --onchip 1 --frames 28416 --steps 1 --warmup 1 --no-configs --no-cpu --no-int16 --no-e2e > gpurun_out/ncu_onchip_v2.log 2>&1
tail -2 gpurun_out/ncu_onchip_v2.log
ls -la gpurun_out | tail -6
