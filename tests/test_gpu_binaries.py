"""GPU: the reference's OWN binaries rebuilt with the GPU stage (BASELINE.json configs 1 and 5).

baseline/build_ref_tree.sh copies the reference tree, applies baseline/apply_gpu_patch.py (--dec-hw GPU in factory::DVBS2, the
rebinding of src/mains/TX_RX_BB/main.cpp:87-91 and src/mains/RX/main_sched.cpp:213-220) and builds
  baseline/_ref/bin/cpu/{dvbs2_tx_rx_bb,dvbs2_tx,dvbs2_ch,dvbs2_rx}   unmodified reference
  baseline/_ref/bin/gpu/{dvbs2_tx_rx_bb,dvbs2_rx}                     the same mains with Decoder_chain_DVBS2_cuda bound in
in the build container; the binaries travel with the repo snapshot.  Here they run on the GPU box:
  * dvbs2_tx_rx_bb (config 1): FER of the GPU binary inside the binomial confidence interval of the CPU binary's, StreamPU in
    the loop, float and sc16 sockets;
  * dvbs2_tx -> dvbs2_ch -> dvbs2_rx on out_tx_noisy.bin (config 5): the GPU receiver decodes the same file as the CPU
    receiver with the same frame / error counts up to demapper rounding, GPU stage replicated by the scheduler.
"""
import math
import os
import re
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
RUN = os.path.join(REF, "run")
CPU, GPU = os.path.join(REF, "bin", "cpu"), os.path.join(REF, "bin", "gpu")
THREADS = min(16, os.cpu_count() or 4)


def _need(*paths):
    for p in paths:
        if not os.path.exists(p):
            pytest.skip("%s was not built (baseline/build_ref_tree.sh needs the reference tree)" % os.path.relpath(p, ROOT))
    os.makedirs(RUN, exist_ok=True)


def _run(cmd, timeout=900):
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.join(ROOT, "dvbs2_b200") + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run(cmd, cwd=RUN, capture_output=True, text=True, timeout=timeout, env=env)
    assert r.returncode == 0, " ".join(cmd) + "\n" + r.stdout[-3000:] + r.stderr[-3000:]
    return r.stdout


def _report(out):
    """last Monte-Carlo report line: Es/N0 | Eb/N0 || FRA | BE | FE | BER | FER || SIM_THR | ET"""
    rows = [ln for ln in out.splitlines() if ln.strip() and not ln.startswith("#") and ln.count("|") >= 8 and "[" not in ln]
    # the terminal may have written intermediate reports on the same line (carriage returns / '*' markers): keep the last one
    f = [x.strip() for x in re.split(r"\|+", rows[-1].split("\r")[-1])]
    return dict(fra=int(f[2]), be=int(f[3]), fe=int(f[4]), fer=float(f[6]), thr=float(f[7]))


def _task_rate(out, name):
    for ln in out.splitlines():
        if name in ln and ln.startswith("#"):
            c = [x.strip() for x in ln.split("|")]
            return float(c[-6]) if len(c) > 6 else None  # AVERAGE throughput column (Mb/s)
    return None


BB = ["--mod-cod", "QPSK-S_8/9", "--dec-implem", "NMS", "--dec-ite", "10", "-m", "4.0", "-M", "4.05", "--ter-freq", "0"]


def test_dvbs2_tx_rx_bb_gpu_binary_matches_cpu_binary_fer():
    _need(os.path.join(CPU, "dvbs2_tx_rx_bb"), os.path.join(GPU, "dvbs2_tx_rx_bb"))
    # the unmodified reference exactly as BASELINE.json config 1 runs it (INTER SIMD, -F 16); thread count capped through the
    # patched binary's --sim-threads would change the binary, so the CPU run uses the CPU binary as is
    cpu = _report(_run([os.path.join(CPU, "dvbs2_tx_rx_bb")] + BB + ["--dec-simd", "INTER", "-F", "16", "-e", "150"]))
    res = {"cpu": cpu}
    for sockets in ("FLOAT", "SC16"):
        out = _run([os.path.join(GPU, "dvbs2_tx_rx_bb")] + BB + ["--dec-hw", "GPU", "--gpu-sockets", sockets, "-F", "2048",
                                                                   "--sim-threads", str(min(4, THREADS)), "-e", "400", "--sim-stats"])
        g = _report(out)
        res[sockets] = dict(g, stage_mbps=_task_rate(out, "demodulate_decode"))
        # the GPU task's sub-timers (device time per stage, Cuda_handle::stage_timers) are rows of the --sim-stats table
        for key in ("demap", "ldpc", "bch", "unpack"):
            assert re.search(r"\|\s+%s \|\|" % key, out), (key, out[-2500:])
        # same channel model, independent noise: the two FER estimates must agree within 4 sigma of their binomial spread
        p = (cpu["fe"] + g["fe"]) / (cpu["fra"] + g["fra"])
        sd = math.sqrt(p * (1 - p) * (1 / cpu["fra"] + 1 / g["fra"]))
        assert abs(cpu["fer"] - g["fer"]) <= 4 * sd + 1e-4, (sockets, cpu, g)
        assert g["fe"] >= 400 and g["fra"] >= 8192
    print("dvbs2_tx_rx_bb:", res)
    with open(os.path.join(ROOT, "gpurun_out", "r2_binaries_tx_rx_bb.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else os.devnull, "w") as f:
        import json

        json.dump(res, f)


def test_dvbs2_rx_gpu_binary_on_out_tx_noisy_bin():
    _need(os.path.join(CPU, "dvbs2_tx"), os.path.join(CPU, "dvbs2_ch"), os.path.join(CPU, "dvbs2_rx"), os.path.join(GPU, "dvbs2_rx"))
    src = ["--src-type", "USER", "--src-path", "../conf/src/K_14232.src", "--mod-cod", "QPSK-S_8/9"]
    for f in ("out_tx.bin", "out_tx_noisy.bin"):
        if os.path.exists(os.path.join(RUN, f)):
            os.remove(os.path.join(RUN, f))
    # README "Testing TX and RX Separately": TX -> file, channel tool adds AWGN at Eb/N0 4.0 dB, RX reads the noisy file
    _run([os.path.join(CPU, "dvbs2_tx"), "--rad-type", "USER_BIN", "--rad-tx-file-path", "out_tx.bin", "-F", "8"] + src +
         ["--tx-time-limit", "700"])
    _run([os.path.join(CPU, "dvbs2_ch"), "--rad-type", "USER_BIN", "--rad-rx-file-path", "out_tx.bin", "--rad-tx-file-path",
          "out_tx_noisy.bin", "--rad-rx-no-loop", "-F", "8", "--mod-cod", "QPSK-S_8/9", "-m", "4.0"])
    n_pl = os.path.getsize(os.path.join(RUN, "out_tx_noisy.bin")) // (2 * 8370 * 2 * 4)  # float32 I/Q, osf 2 (Radio_user_binary.cpp:74-97)
    assert n_pl >= 512, n_pl
    rx = ["--sim-stats"] + src + ["--rad-type", "USER_BIN", "--rad-rx-file-path", "out_tx_noisy.bin", "--dec-implem", "NMS",
                                  "--dec-ite", "10", "--snk-path", "/dev/null", "--rad-rx-no-loop", "--no-wl-phases", "-R",
                                  str(THREADS)]
    c = _report(_run([os.path.join(CPU, "dvbs2_rx")] + rx + ["-F", "16", "--dec-simd", "INTER"]))
    res = {"pl_frames_in_file": n_pl, "cpu": c}
    for sockets in ("FLOAT", "SC16"):
        out = _run([os.path.join(GPU, "dvbs2_rx")] + rx + ["-F", "16", "--dec-hw", "GPU", "--gpu-sockets", sockets])
        g = _report(out)
        res[sockets] = g
        # same file, same synchronisers: the same frames reach the decode stage; only demapper rounding (and, for SC16, the
        # 16-bit samples) may move a marginal frame across the failure line
        assert g["fra"] == c["fra"], (c, g)
        assert abs(g["fe"] - c["fe"]) <= 2 + c["fe"] // 10, (c, g)
        assert "Decoder_chain" in out  # the GPU stage is a task of the scheduled pipeline
    print("dvbs2_rx:", res)
    with open(os.path.join(ROOT, "gpurun_out", "r2_binaries_rx.json") if os.path.isdir(os.path.join(ROOT, "gpurun_out"))
              else os.devnull, "w") as f:
        import json

        json.dump(res, f)
    for f in ("out_tx.bin", "out_tx_noisy.bin"):
        os.remove(os.path.join(RUN, f))
