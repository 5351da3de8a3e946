#!/usr/bin/env bash
# round 2, second GPU pass: the reference's own binaries with the GPU stage, the C++ modules, compact sockets, demapper, bench
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
timeout -k 10 1500 python -m pytest tests/test_gpu_binaries.py -q -x -s 2>&1 | tail -25 | tee gpurun_out/r2_tests_binaries.log
timeout -k 10 1500 python -m pytest tests/test_gpu_cpp_modules.py tests/test_gpu_compact_sockets.py tests/test_gpu_bch_demod_chain.py tests/test_gpu_rx_front.py tests/test_gpu_golden.py -q -x 2>&1 | tail -8 | tee gpurun_out/r2_tests_b.log
timeout 1500 python bench.py --no-configs 2> gpurun_out/r2_bench_b.err | tee gpurun_out/r2_bench_b.json | cut -c1-300
tail -5 gpurun_out/r2_bench_b.err
for nc in 2 4; do timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-clones $nc --e2e-sockets sc16_packed,sc8_packed --steps 6 2>/dev/null > gpurun_out/r2_bench_b_nc$nc.json; done
timeout 600 python bench.py --no-configs --no-cpu --no-int16 --e2e-clones 1 --e2e-chunks 0 --e2e-frames 340992 --e2e-sockets sc16_packed --steps 6 2>/dev/null > gpurun_out/r2_bench_b_nc1.json
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_b*.json")):
    try:
        d = json.load(open(f))
        e = d["e2e"]
        print(f, d["value"], d["stage_ms"], "e2e", {k: (v.get("value"), v.get("ms_per_step")) for k, v in e["sockets"].items()}, "int16", (d.get("int16") or {}).get("value"), ((d.get("int16") or {}).get("e2e") or {}).get("value"))
    except Exception as ex:
        print(f, "failed", ex)
PY
