#!/usr/bin/env python
"""Benchmark of the DVB-S2 baseband decode chain (demap -> LDPC NMS-10 -> BCH) on B200 — driver contract.

    python bench.py [--gpus N --steps K --warmup W] [--impl reference] [--workload qpsk89|qpsk35|8psk35|8psk89|16apsk89]

One "step" = one pass of the fused chain (dvbs2_cuda_chain_dev) over one batch of F synthetic AWGN frames per GPU.
`value`  = info Mbit/s (K_bch * frames / time, the reference's throughput definition, D/mains/TX_RX_BB/main.cpp:130) with
           the received symbols already resident in HBM; whole-job aggregate over all ranks (frames shard by batch, no
           collective on the data path; NCCL only reduces the time and the FER counters).
`e2e`    = the same metric through the host-buffer C-ABI calls with pinned HOST buffers, H2D of the symbols and D2H of the
           payload inside the timed region, driven like a replicated StreamPU stage (module clones on host threads).  Measured
           per socket format (`e2e.sockets`): the reference's own socket types (float I/Q in, int32 per bit out: 121.7 KB per
           frame, PCIe-bound near 11 Gbit/s on any GPU) and the compact formats of SURVEY §8f N2 (sc16 / sc8 I/Q in, packed
           bytes out).  `e2e.value` is the sc16-in / packed-out stage: bit-identical to the float sockets for the same
           sample values (tests/test_gpu_compact_sockets.py), i.e. what a radio that delivers sc16 feeds the stage.
`roofline` = the LDPC kernel (dominant), algorithmic bytes per frame from SURVEY.md §8(d) / DESIGN.md §5.
`cpu_baseline` = the compiled reference chain (oracle/_ref, AVX-512 ONMS_inter -F 16 + Modem_generic_fast + BCH_std) on
           this box's host cores, bounded sample; `--impl reference` times just that.
Multi-GPU: launched by torchrun (one rank per GPU); barrier + synchronize around the timed region, MAX over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json metric/config[0]: QPSK-S_8/9 at Eb/N0 4.0 dB;  config[1]: QPSK-S_3/5 at the NMS-10 waterfall
    "qpsk89": dict(modcod="QPSK-S_8/9", ebn0=4.0),
    "qpsk35": dict(modcod="QPSK-S_3/5", ebn0=2.2),
    "8psk35": dict(modcod="8PSK-S_3/5", ebn0=3.9),
    "8psk89": dict(modcod="8PSK-S_8/9", ebn0=7.5),
    "16apsk89": dict(modcod="16APSK-S_8/9", ebn0=7.6),
}
N_LDPC = 16200
HBM_FALLBACK_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--workload", default="qpsk89", choices=sorted(WORKLOADS))
    p.add_argument("--ebn0", type=float, default=None)
    p.add_argument("--prec", default="f32", choices=["f32", "i16", "i8"])
    p.add_argument("--ite", type=int, default=10)
    p.add_argument("--frames", type=int, default=0, help="frames per GPU per step (default: the largest of 3, 2, 1 waves of the LDPC kernel = SMs x 12 warps x 64 frames each that fits in 80 %% of the free HBM)")
    p.add_argument("--e2e-frames", type=int, default=0, help="frames per e2e call of one module clone (default: one wave of the LDPC kernel)")
    p.add_argument("--e2e-clones", type=int, default=0, help="module clones (handles + host threads) of the e2e run; 0 = automatic: 4 x one wave per call for the float sockets (3: 0.80-0.83 of the device-resident rate, 4-5: 0.85, 6 does not fit in HBM), 2 x the whole batch per call for the compact sockets")
    p.add_argument("--e2e-chunks", type=int, default=-1, help="host_pipeline_chunks of the e2e clones (-1: 1 when there are several clones, else automatic)")
    p.add_argument("--e2e-sockets", default="", help="comma list of socket formats to time (default: f32_int32,f32_packed,sc16_packed,sc8_packed; llr8_packed with --prec i16)")
    p.add_argument("--no-configs", action="store_true", help="skip the extra lines for BASELINE configs 2-4")
    p.add_argument("--no-numa", action="store_true", help="do not bind this rank to the CPU cores next to its GPU")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--no-cpu", action="store_true")
    p.add_argument("--no-int16", action="store_true", help="skip the extra fixed-point (int16) line of the same workload")
    p.add_argument("--no-compaction", action="store_true")
    p.add_argument("--rounds", type=int, default=0, help="decoding rounds (frame compaction between rounds); 0 = library default")
    p.add_argument("--exit-eighths", type=int, default=0, help="hand-over threshold in eighths of a group; 0 = library default")
    p.add_argument("--nf-later", type=int, default=0, help="f32: frames per lane in rounds >= 2; 0 = library default")
    p.add_argument("--nf-first", type=int, default=0, help="f32: frames per lane in round 1 (1 = 32-frame groups, 2 = 64-frame groups); 0 = library default (by batch size)")
    p.add_argument("--coop-max", type=int, default=0, help="f32: largest call that uses the cooperative small-batch LDPC kernel (0 = library default 16384, -1 = never)")
    p.add_argument("--onchip", type=int, default=0, help="f32: 1 = the on-chip LDPC kernel (design B), -1 = never, 0 = library default")
    p.add_argument("--cpu-seconds", type=float, default=12.0)
    p.add_argument("--cpu-threads", type=int, default=0)
    return p.parse_args()


def hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            d = json.load(f)
        for k in ("hbm_gbs", "hbm_gb_s", "hbm_copy_gbs"):
            if k in d:
                return float(d[k]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons of one GPU every 200 ms while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        rows = [r for t, r in self.rows if t0 - 0.1 <= t <= t1 + 0.3 and len(r) >= 7] or [r for _, r in self.rows if len(r) >= 7]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(int(float(r[0])) for r in rows)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[3 + i].lower().startswith("active") for r in rows)]
        pw = [float(r[2]) for r in rows if r[2].replace(".", "", 1).isdigit()]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": int(float(rows[0][1])), "reasons": reasons,
                "samples": len(rows), "power_w_max": max(pw) if pw else None}


def host_threads(args):
    if args.cpu_threads > 0:
        return args.cpu_threads
    n = os.cpu_count() or 1
    try:  # every reference BCH clone allocates a 1.07 GB table
        import psutil

        n = max(1, min(n, int(psutil.virtual_memory().available * 0.6 / 1.2e9)))
    except Exception:
        pass
    return n


def numa_bind(gpu_index):
    """Bind this process (and so its pinned staging buffers, first touch) to the CPU cores next to its GPU: with one rank per
    GPU the ranks' host traffic then spreads over all sockets' memory controllers instead of piling up on node 0."""
    try:
        import pynvml

        pynvml.nvmlInit()
        hd = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
        n = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(hd, n)
        cpus = {64 * i + b for i, w in enumerate(mask) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return {"bound": False, "why": "empty affinity mask"}
        os.sched_setaffinity(0, cpus)
        node = None
        for d in sorted(os.listdir("/sys/devices/system/node")) if os.path.isdir("/sys/devices/system/node") else []:
            if d.startswith("node") and os.path.exists("/sys/devices/system/node/%s/cpu%d" % (d, min(cpus))):
                node = int(d[4:])
        return {"bound": True, "cpus": len(cpus), "numa_node": node}
    except Exception as e:  # noqa: BLE001
        return {"bound": False, "why": repr(e)[:120]}


def cpu_reference_run(modcod, ite, sigma, sym, seconds, threads):
    """Frames/s of the compiled reference chain (demodulate + deinterleave + ONMS_inter LDPC + BCH) on `threads` cores."""
    import numpy as np

    from oracle.oracle import N_LDPC as N, Ref

    if not Ref.available():
        return None
    r = Ref()
    F = sym.shape[0]
    llr = np.zeros((F, N), np.float32)
    fps, done = r.bench_chain(modcod, ite, threads, seconds, 7, sigma, sym, llr)
    r.close()
    return fps, done


def host_sample(modcod, ebn0, n_frames, seed=0xD5B52):
    """Small host batch of the bench workload for the CPU arm: random payloads through the TX chain + AWGN (oracle's
    synth_frames — random codewords matter on the CPU: the reference BCH skips zero bits in its syndrome loop)."""
    from oracle.oracle import Oracle, synth_frames

    s = synth_frames(Oracle(), modcod, n_frames, ebn0, seed=seed)
    return s["sigma"], s["rx"]


def run_reference(args):
    """--impl reference: the reference's own CPU implementation of the path, all host threads, bounded sample/step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from dvbs2_b200 import synth

    wl = WORKLOADS[args.workload]
    modcod, ebn0 = wl["modcod"], wl["ebn0"] if args.ebn0 is None else args.ebn0
    K_bch = synth.MODCODS[modcod]["K_bch"]
    try:
        from oracle.oracle import Ref

        ok = Ref.available()
    except Exception as e:  # noqa: BLE001
        ok = False
    if not ok:
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdvbs2_ref.so not built on this host"}))
        return
    threads = host_threads(args)
    sig, rx = host_sample(modcod, ebn0, 512)
    n_calls = args.steps + args.warmup
    sec = max(0.5, min(4.0, 100.0 / max(1, n_calls)))
    fps = []
    for i in range(n_calls):
        r = cpu_reference_run(modcod, args.ite, sig, rx, sec, threads)
        if i >= args.warmup:
            fps.append(r[0])
    v = K_bch * (sum(fps) / len(fps)) / 1e6
    sample = "%d x %.1f s of the chain over 512 frames cycling, %d threads x ONMS_inter -F 16" % (len(fps), sec, threads)
    print(json.dumps({
        "impl": "reference", "metric": "info Mbit/s, %s NMS %d-iter decode chain (demap+LDPC+BCH)" % (modcod, args.ite),
        "value": round(v, 3), "unit": "Mbit/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(sec * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (random payloads through the TX chain + AWGN, host-generated)",
        "config": {"workload": "%s Eb/N0 %.2f dB NMS norm 1.0 %d it, early stop" % (modcod, ebn0, args.ite)},
        "cpu_baseline": {"value": round(v, 3), "unit": "Mbit/s", "cores": threads, "kind": "reference", "sample": sample},
        "e2e": {"value": round(v, 3), "unit": "Mbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    T0 = time.time()
    phases = {}  # wall-clock seconds of each part of this run (reported as "wall_s": how long the default command takes)

    def lap(name, since):
        phases[name] = round(phases.get(name, 0.0) + time.time() - since, 1)
        return time.time()

    import numpy as np
    import torch
    import torch.distributed as dist

    from dvbs2_b200 import (Decoder_chain_compact_DVBS2_cuda, Decoder_chain_DVBS2_cuda, Decoder_llr8_DVBS2_cuda, Handle,
                            Modem_DVBS2_cuda, Transmitter_DVBS2_cuda, synth)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: there is no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = {"bound": False, "why": "--no-numa"} if args.no_numa else numa_bind(local)
    real_stdout = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its banner ("NCCL version ...") with printf on the process's stdout when the image sets NCCL_DEBUG; the
        # contract is ONE JSON line there, so fd 1 points at stderr until the line is printed
        sys.stdout.flush()
        real_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local])

    def reduce_(x, op):
        t = torch.tensor(x, dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=op)
        return t.tolist()

    wl = WORKLOADS[args.workload]
    modcod, ebn0 = wl["modcod"], wl["ebn0"] if args.ebn0 is None else args.ebn0
    mc = synth.MODCODS[modcod]
    K_bch, N_bch, bps = mc["K_bch"], mc["N_bch"], mc["bps"]
    n_sym = N_LDPC // bps
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    NB = 2  # input batches rotated between steps (each is far larger than the 126 MB L2)
    # frames one launch of the LDPC kernel holds in flight: 12 warps per SM, 64 (f32) or 128 (i16, i8) frames per warp
    wave = sms * 12 * (64 if args.prec == "f32" else 128)
    if args.frames > 0:
        F = args.frames
    else:
        # several waves per launch let the work queue balance groups that stop at different iterations (measured:
        # 17.4 / 18.3 / 19.2 Gbit/s at 1 / 2 / 3 waves, profiles/README.md); bounded by what fits in HBM beside the inputs
        esz0 = {"f32": 4, "i16": 2, "i8": 1}[args.prec]
        # decoder state incl. the compaction regions (laid out for a hand-over threshold of 6/8 float, 7/8 int16) and the later rounds' slots
        state = {"f32": 1.85, "i16": 2.1, "i8": 1.6}[args.prec]
        cn_b = (12 if N_bch == 9720 else 16) // (2 if args.prec == "i16" else 1)  # compressed check state per check and frame
        per_frame = state * (N_LDPC * esz0 + (N_LDPC - N_bch) * cn_b + N_LDPC // 4) + NB * (2 * n_sym * 4 + 4) + K_bch * 4 + 16
        free_b = torch.cuda.mem_get_info(dev)[0]
        F = next((k * wave for k in (3, 2) if k * wave * per_frame < 0.8 * free_b), wave)
        if world > 1:  # every rank decodes the same number of frames per step (weak scaling): agree on the smallest
            t = torch.tensor([F], dtype=torch.int64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            F = int(t.item())

    tune = dict(compaction=not args.no_compaction, compaction_rounds=args.rounds, compaction_exit_eighths=args.exit_eighths,
                compaction_nf_later=args.nf_later, ldpc_coop_max_frames=args.coop_max, ldpc_onchip=args.onchip,
                ldpc_frames_per_lane=args.nf_first)
    h = Handle(modcod=modcod, n_ite=args.ite, norm=1.0, precision=args.prec, early_stop=True, n_frames=F, device=local, **tune)
    chain = Decoder_chain_DVBS2_cuda(h, F)
    # synthetic frames made where they are decoded: random payload -> BB scramble -> BCH -> LDPC -> interleave -> map -> AWGN
    tx = Transmitter_DVBS2_cuda(h, F)
    sigma = synth.ebn0_to_sigma(modcod, ebn0)
    seeds = [0xD5B52 + 1000 * rank + b for b in range(NB)]
    batches = []
    for sd in seeds:
        sig = torch.empty(F, dtype=torch.float32, device=dev)
        rx = torch.empty((F, 2 * n_sym), dtype=torch.float32, device=dev)
        tx.random_payload(sd)
        tx.modulate_awgn_dev(sigma, sd, sig, rx)
        batches.append((sig, rx))
    h.sync()
    V = torch.zeros((F, K_bch), dtype=torch.int32, device=dev)
    c1 = torch.zeros(F, dtype=torch.int8, device=dev)
    c2 = torch.zeros(F, dtype=torch.int8, device=dev)
    nit = torch.zeros(F, dtype=torch.int32, device=dev)
    ext = torch.cuda.ExternalStream(h.stream(), device=dev)

    def step(i):
        sig, rx = batches[i % NB]
        chain.demodulate_decode_dev(sig, rx, c1, c2, V, nit)

    torch.cuda.synchronize()
    tp = lap("setup_and_tx", T0)
    for i in range(args.warmup):
        step(i)
    h.sync()
    h.reset_stats()
    h.set_profiling(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler = ClockSampler(local)
    barrier()
    torch.cuda.synchronize()
    t0 = time.time()
    with torch.cuda.stream(ext):
        e0.record()
    for i in range(args.steps):
        step(i)
    with torch.cuda.stream(ext):
        e1.record()
    h.sync()
    torch.cuda.synchronize()
    t1 = time.time()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop(t0, t1)
    tm = h.timers()
    st = h.stats()
    h.set_profiling(False)
    ms_max = reduce_([ms], dist.ReduceOp.MAX if world > 1 else None)[0]
    # FER / iteration counters of the timed region: NCCL sum over ranks (the Monitor_reduction replacement)
    cnt = reduce_([st["frames"], st["ldpc_iterations"], st["ldpc_failures"], st["bch_failures"], st["bch_nonzero"]],
                  dist.ReduceOp.SUM if world > 1 else None)
    frames_total = F * args.steps * world
    value = K_bch * frames_total / (ms_max * 1e-3) / 1e6
    # untimed scoring pass: decode each batch once more and compare with its payload on the device (BFER monitor)
    h.reset_stats()
    for b, sd in enumerate(seeds):
        tx.random_payload(sd)  # same seed -> same payload words as when the batch was made
        step(b)
        tx.check_errors()
    mon = tx.monitor()
    mon = reduce_([mon["n_frames"], mon["n_be"], mon["n_fe"]], dist.ReduceOp.SUM if world > 1 else None)

    tp = lap("warmup_timed_scoring", tp)
    # ---- roofline of the dominant kernel (LDPC decode), live CUDA-event time per launch
    peak, peak_src = hbm_peak()
    ldpc_ms = tm["ldpc_ms"] / max(1, tm["ldpc_calls"])
    esz = {"f32": 4, "i16": 2, "i8": 1}[args.prec]
    alg_bytes_frame = esz * N_LDPC + 4 * N_bch + 1  # compulsory I/O of decode_siho with AFF3CT socket types
    achieved = alg_bytes_frame * F / (ldpc_ms * 1e-3) / 1e9
    M = N_LDPC - N_bch
    d_info = 25 if N_bch == 14400 else 9
    E_info = M * d_info
    it_per_frame = st["ldpc_iterations"] / max(1, st["frames"])
    # check state read + written per check and frame (packed pairs for int16; 12-byte state for the degree-11 checks of rate 3/5)
    cn_rw = 2 * (12 if d_info == 9 else 16) // (2 if args.prec == "i16" else 1)
    msg_bytes_frame = it_per_frame * (2 * esz * E_info + 2 * esz * M + cn_rw * M)  # DESIGN.md §5 message-passing traffic
    roofline = {"bound": "hbm", "kernel": "ldpc_decode_kernel", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 5), "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": alg_bytes_frame, "kernel_ms": round(ldpc_ms, 4),
                "kernel_share_of_step": round(tm["ldpc_ms"] / ms, 4),
                "message_passing": {"bytes_per_frame": round(msg_bytes_frame), "achieved": round(msg_bytes_frame * F / (ldpc_ms * 1e-3) / 1e9, 1),
                                    "frac": round(msg_bytes_frame * F / (ldpc_ms * 1e-3) / 1e9 / peak, 4)}}
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            per_frame = json.load(f).get("per_frame", {}).get(args.workload + "_" + args.prec)
            roofline["traffic"] = None if per_frame is None else int(per_frame * F)  # DRAM bytes of the LDPC launches of a step
    except Exception:
        pass
    stage_ms = {k[:-3]: round(tm[k] / max(1, tm[k[:-3] + "_calls"]), 4) for k in tm if k.endswith("_ms") and tm[k[:-3] + "_calls"]}
    launches = st["kernel_launches"]

    # ---- end to end through the host-buffer C ABI (pinned host memory, H2D + D2H inside the timed region).
    # Driven the way StreamPU drives a replicated stage: NC module clones (one handle + streams each), one host thread
    # per clone calling the synchronous task; copies of one clone overlap the kernels of the others.  One run per socket
    # format; the clones of a format are created, timed and destroyed before the next format allocates its buffers.
    e2e = None
    if not args.no_e2e:
        args.e2e_clones_given = args.e2e_clones > 0
        NC0 = args.e2e_clones if args.e2e_clones > 0 else 4
        # device-resident result of the first frames of batch 0 (what clone 0 decodes), then the big decoder state goes: the e2e
        # clones bring their own handles
        step(0)
        h.sync()
        V_head = V[:min(F, wave)].cpu().numpy()
        h.close()
        h = None
        del V, chain, tx
        # batch 0 moves to host memory (the e2e inputs live there anyway) and every device batch goes: the clones of the compact
        # formats then have room for three-wave calls
        host0 = (batches[0][0].cpu(), batches[0][1].cpu())
        batches = None
        sig = rx = None  # the loop variables of the batch generation still held the last batch
        torch.cuda.empty_cache()
        default_fmts = "llr8_packed" if args.prec == "i16" else "f32_int32,f32_packed,sc16_packed,sc12_packed,sc8_packed"
        fmts = [f for f in (args.e2e_sockets or default_fmts).split(",") if f]
        SC16, SC12, SC8 = 1.0 / 8192, 1.0 / 512, 1.0 / 32  # LSB of the integer I/Q formats: +-4.0 full scale (unit-energy constellations)
        pin = lambda *shape, dtype: torch.empty(shape, dtype=dtype).pin_memory()  # noqa: E731

        def e2e_run(fmt, allow_big=True):
            in_kind, out_kind = fmt.split("_")
            # one call of a clone = one wave of the LDPC kernel (every launch fills the machine; the copies of one clone overlap
            # the kernels of the others, so calls are not chunked).  The reference's int32-per-bit sockets need 14 GB of pinned
            # memory per wave and clone: half a wave there when eight ranks share one host.
            # Call shape.  NC clones x one wave per call, except for 8-bit samples: TWO clones x the whole device-resident batch
            # per call — every LDPC launch then has the three waves that make the device-resident figure (single-wave launches lose
            # 15 % to their tails: 16.7 against 19.7 Gbit/s).  That only pays while the H2D copy of a call is short: the copy engine
            # gets 20-35 GB/s, not the link's 55, while the LDPC kernel of the other clone saturates HBM (measured,
            # scripts/gpu_r2_sc12.sh: three-wave calls give 14.4 / 15.6 / 17.8 Gbit/s with sc16 / sc12 / sc8 against 15.9 / - / 15.5
            # for one-wave calls).  --e2e-frames / --e2e-clones given explicitly pin the shape.
            NC = NC0
            hpf = {"f32": 195e3, "i16": 118e3, "i8": 80e3}["i16" if in_kind == "llr8" else args.prec]  # handle bytes per frame (measured)
            big = allow_big and in_kind == "sc8" and args.e2e_frames == 0 and not args.e2e_clones_given
            if big:
                need = 2 * F * (hpf + ({"sc16": 4, "sc12": 3, "sc8": 2}.get(in_kind, 0) * n_sym or N_LDPC) + K_bch // 8 + 64)
                big = need < 0.97 * torch.cuda.mem_get_info(dev)[0]
            if big:
                NC, Fe = 2, F
            else:
                Fe = args.e2e_frames if args.e2e_frames > 0 else (wave // 2 if (fmt == "f32_int32" and world > 1) else wave)
                Fe = max(256, min(F if args.e2e_frames > 0 else F // NC, Fe) // 256 * 256)
            # page-locked host memory of all clones must fit beside everything else in this box's RAM (shared by all ranks)
            in_kind0, out_kind0 = fmt.split("_")
            per_frame_host = ({"f32": 8, "sc16": 4, "sc12": 3, "sc8": 2}.get(in_kind0, 0) * n_sym or N_LDPC) + (4 * K_bch if out_kind0 == "int32" else K_bch // 8) + 16
            try:
                import psutil

                budget = 0.5 * psutil.virtual_memory().available / max(1, world)
                if NC * Fe * per_frame_host > budget:
                    Fe = max(8192, int(budget / (NC * per_frame_host)) // 256 * 256)
            except ImportError:
                pass
            if world > 1:  # the value is computed from one shape: every rank uses the smallest
                t = torch.tensor([Fe, NC], dtype=torch.int64, device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MIN)
                Fe, NC = int(t[0].item()), int(t[1].item())
            prec = "i16" if in_kind == "llr8" else args.prec
            chunks = args.e2e_chunks if args.e2e_chunks >= 0 else (1 if NC > 1 else 0)
            clones, same = [], None
            # ---- phase 1: the host (socket) buffers of every clone, made with device temporaries that are gone before the handles come
            host_bufs = []
            for c in range(NC):
                o = (c * Fe) % max(1, F - Fe + 1)  # clones take disjoint slices of batch 0 while they fit
                hs = pin(Fe, dtype=torch.float32)
                hs.copy_(host0[0][o:o + Fe])
                hc1, hc2, hn = pin(Fe, dtype=torch.int8), pin(Fe, dtype=torch.int8), pin(Fe, dtype=torch.int32)
                if in_kind == "f32":
                    hx = pin(Fe, 2 * n_sym, dtype=torch.float32)
                    hx.copy_(host0[1][o:o + Fe])
                elif in_kind in ("sc16", "sc12", "sc8"):  # the radio's ADC words: quantised once, outside the timed region
                    sc, lim, dt = {"sc16": (SC16, 32767, torch.int16), "sc12": (SC12, 2047, torch.int32), "sc8": (SC8, 127, torch.int8)}[in_kind]
                    hx = pin(Fe, 3 * n_sym, dtype=torch.uint8) if in_kind == "sc12" else pin(Fe, 2 * n_sym, dtype=dt)
                    for a0 in range(0, Fe, wave):  # a wave at a time through the device
                        q = host0[1][o + a0:o + min(Fe, a0 + wave)].to(dev) * (1.0 / sc)
                        q = (torch.sign(q) * torch.floor(torch.abs(q) + 0.5)).clamp_(-lim, lim).to(dt)
                        if in_kind == "sc12":  # I = bits 0..11, Q = bits 12..23 of a little-endian 24-bit word per symbol
                            w24 = (q[:, 0::2] & 0xfff) | ((q[:, 1::2] & 0xfff) << 12)
                            q = torch.stack((w24 & 0xff, (w24 >> 8) & 0xff, (w24 >> 16) & 0xff), dim=-1).to(torch.uint8).reshape(q.shape[0], -1)
                            del w24
                        hx[a0:a0 + q.shape[0]].copy_(q)
                        del q
                else:  # llr8: Quantizer_pow2(6, 2) outputs of the demapper's LLRs, one int8 per LLR
                    hx = pin(Fe, N_LDPC, dtype=torch.int8)
                    Fs = min(Fe, 113664)
                    hf = Handle(modcod=modcod, n_frames=Fs, device=local)
                    llr = torch.empty((Fs, N_LDPC), dtype=torch.float32, device=dev)
                    for a0 in range(0, Fe, Fs):
                        n = min(Fs, Fe - a0)
                        ssig, srx = host0[0][o + a0:o + a0 + n].to(dev), host0[1][o + a0:o + a0 + n].to(dev)
                        torch.cuda.synchronize()
                        Modem_DVBS2_cuda(hf, n).demodulate_dev(ssig, srx, llr[:n], deinterleave=True)
                        hf.sync()
                        q = llr[:n] * 4.0
                        hx[a0:a0 + n].copy_((torch.sign(q) * torch.floor(torch.abs(q) + 0.5)).clamp_(-31, 31).to(torch.int8))
                        del q, ssig, srx
                    hf.close()
                    del llr
                hV = pin(Fe, K_bch, dtype=torch.int32) if out_kind == "int32" else pin(Fe, K_bch // 8, dtype=torch.uint8)
                host_bufs.append((hs, hx, hV, hc1, hc2, hn))
            torch.cuda.empty_cache()
            # ---- phase 2: one handle per clone
            gpu_free0 = torch.cuda.mem_get_info(dev)[0]
            for c in range(NC):
                try:
                    hc = Handle(modcod=modcod, n_ite=args.ite, norm=1.0, precision=prec, early_stop=True, n_frames=Fe, device=local,
                                host_pipeline_chunks=chunks, **tune)
                except Exception as e:  # noqa: BLE001
                    for old in clones:
                        old[0].close()
                    if big:  # the estimate was too tight for this box: the one-wave shape instead
                        del clones, host_bufs
                        torch.cuda.empty_cache()
                        return e2e_run(fmt, allow_big=False)
                    raise RuntimeError("clone %d of %d x %d frames: %r (GPU memory free before the clones %.1f GB, now %.1f GB)"
                                       % (c, NC, Fe, e, gpu_free0 / 1e9, torch.cuda.mem_get_info(dev)[0] / 1e9))
                hs, hx, hV, hc1, hc2, hn = host_bufs[c]
                a = dict(sig=hs.numpy(), x=hx.numpy(), c1=hc1.numpy(), c2=hc2.numpy(), n=hn.numpy(), V=hV.numpy())
                if in_kind == "f32":
                    m = Decoder_chain_DVBS2_cuda(hc, Fe)
                    call = ((lambda m=m, a=a: m.demodulate_decode(a["sig"], a["x"], a["c1"], a["c2"], a["V"], a["n"]))
                            if out_kind == "int32" else
                            (lambda m=m, a=a: m.demodulate_decode_packed(a["sig"], a["x"], a["c1"], a["c2"], a["V"], a["n"], descramble=True)))
                elif in_kind == "llr8":
                    m = Decoder_llr8_DVBS2_cuda(hc, Fe)
                    call = lambda m=m, a=a: m.decode(a["x"], a["c1"], a["c2"], V_K_packed=a["V"], n_iter=a["n"], descramble=True)  # noqa: E731
                else:
                    m = Decoder_chain_compact_DVBS2_cuda(hc, Fe, in_kind, {"sc16": SC16, "sc12": SC12, "sc8": SC8}[in_kind])
                    call = lambda m=m, a=a: m.demodulate_decode(a["sig"], a["x"], a["c1"], a["c2"], V_K_packed=a["V"], n_iter=a["n"], descramble=True)  # noqa: E731
                clones.append((hc, call, a, host_bufs[c]))
            torch.cuda.synchronize()

            # one call of clone 0 alone: warm-up, and the length of a call when nothing shares the GPU or the link
            t_solo = time.perf_counter()
            clones[0][1]()
            t_solo = time.perf_counter() - t_solo

            def run_clones(n):
                errors = []

                def work(c):
                    try:
                        # clones start a fraction of a call apart, as the clones of a replicated StreamPU stage do (each is handed
                        # its batch when the producer has filled it): the copies of one clone then fall under the kernels of the
                        # others instead of all clones copying, then all computing, in step
                        time.sleep(c * t_solo / NC)
                        for _ in range(n):
                            clones[c][1]()
                    except BaseException as e:  # noqa: BLE001 — a clone that stopped early must void the whole measurement
                        errors.append(e)
                ths = [threading.Thread(target=work, args=(c,)) for c in range(NC)]
                for t in ths:
                    t.start()
                for t in ths:
                    t.join()
                if errors:
                    raise errors[0]

            run_clones(max(1, min(args.warmup, 2)))
            for hc, _, _, _ in clones:
                hc.reset_stats()
            clones[0][0].set_profiling(True)  # device time per stage of clone 0's calls while the other clones share the GPU
            barrier()
            torch.cuda.synchronize()
            w0 = time.perf_counter()
            run_clones(args.steps)
            torch.cuda.synchronize()
            w = time.perf_counter() - w0
            w = reduce_([w], dist.ReduceOp.MAX if world > 1 else None)[0]
            barrier()
            n_launch = sum(hc.stats()["kernel_launches"] for hc, _, _, _ in clones)
            tm0 = clones[0][0].timers()
            stage0 = {k[:-3]: round(tm0[k] / max(1, tm0[k[:-3] + "_calls"]), 3) for k in tm0 if k.endswith("_ms") and tm0[k[:-3] + "_calls"]}
            fails = sum(int((a["c2"] == 0).sum()) for _, _, a, _ in clones)
            if fmt == "f32_int32":  # the host result of clone 0 equals the device-resident result of the same frames
                same = bool(np.array_equal(clones[0][2]["V"][:len(V_head)], V_head[:Fe]))
            in_b = clones[0][2]["x"].nbytes + clones[0][2]["sig"].nbytes
            out_b = clones[0][2]["V"].nbytes + 2 * Fe + 4 * Fe
            for hc, _, _, _ in clones:
                hc.close()
            del clones
            r = {"value": round(K_bch * Fe * NC * args.steps * world / w / 1e6, 2), "unit": "Mbit/s",
                 "h2d_bytes_per_step": int(NC * in_b), "d2h_bytes_per_step": int(NC * out_b),
                 "bytes_per_frame": {"h2d": round(in_b / Fe, 1), "d2h": round(out_b / Fe, 1)},
                 "host_gbs_per_gpu": round(NC * (in_b + out_b) * args.steps / w / 1e9, 2),
                 "host_gbs_total": round(world * NC * (in_b + out_b) * args.steps / w / 1e9, 2),
                 "frames_per_step_per_gpu": Fe * NC, "clones": NC, "ms_per_step": round(w / args.steps * 1e3, 3),
                 "gpu_launches": int(n_launch), "bch_fer": round(fails / (Fe * NC), 5), "clone0_stage_ms_per_call": stage0}
            if same is not None:
                r["matches_device_path"] = same
            return r

        API = {"f32_int32": "dvbs2_cuda_chain: float I/Q in, int32 per bit out (the reference's socket types)",
               "f32_packed": "dvbs2_cuda_chain_packed: float I/Q in, descrambled payload bytes out",
               "sc16_packed": "dvbs2_cuda_chain_compact: sc16 I/Q in (LSB 2^-13), descrambled payload bytes out",
               "sc12_packed": "dvbs2_cuda_chain_compact: sc12 I/Q in (12-bit samples, 3 bytes per symbol, LSB 2^-9), descrambled payload bytes out",
               "sc8_packed": "dvbs2_cuda_chain_compact: sc8 I/Q in (LSB 2^-5), descrambled payload bytes out",
               "llr8_packed": "dvbs2_cuda_decode_llr8: int8 LLRs in (Quantizer_pow2 6/2), int16 decoder, payload bytes out"}
        socks = {}
        for fmt in fmts:
            try:
                socks[fmt] = e2e_run(fmt)
                socks[fmt]["api"] = API.get(fmt, fmt)
            except Exception as e:  # noqa: BLE001
                socks[fmt] = {"value": None, "error": repr(e)[:400], "gpu_free_gb_after_failure": round(torch.cuda.mem_get_info(dev)[0] / 1e9, 1)}
            torch.cuda.empty_cache()
        # headline sockets: the faster of the two integer sample formats whose FER equals the float sockets' inside the Monte-Carlo
        # spread (16- and 12-bit samples, both UHD wire formats); 8-bit samples cost ~0.02 dB and are reported beside it only
        radio = sorted((f for f in ("sc16_packed", "sc12_packed") if socks.get(f, {}).get("value")), key=lambda f: -socks[f]["value"])
        head = next((f for f in radio + ["llr8_packed", "f32_packed", "f32_int32"] if socks.get(f, {}).get("value")), None)
        if head:
            e2e = dict(socks[head])
            e2e["headline_sockets"] = head
            e2e["api"] = API.get(head, head) + "; %d concurrent module clones (host threads) x %d frames per call, one chunk per call" % (
                e2e["clones"], e2e["frames_per_step_per_gpu"] // e2e["clones"])
            e2e["vs_device_resident"] = round(e2e["value"] / value, 4)
            e2e["host_numa"] = numa
            e2e["sockets"] = socks
            if "f32_int32" in socks:
                e2e["matches_device_path"] = socks["f32_int32"].get("matches_device_path")
        else:  # every format failed: keep the reasons
            e2e = {"value": None, "sockets": socks}

    tp = lap("e2e", tp)
    # ---- CPU baseline beside it (rank 0, N=1 only): the compiled reference on this box's host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        try:
            threads = host_threads(args)
            sig, rx = host_sample(modcod, ebn0, 512)
            r = cpu_reference_run(modcod, args.ite, sig, rx, args.cpu_seconds, threads)
            if r is not None:
                cpu = {"value": round(K_bch * r[0] / 1e6, 3), "unit": "Mbit/s", "cores": threads, "kind": "reference",
                       "sample": "%.0f s of demap+deinterleave+ONMS_inter(-F 16)+BCH_std over 512 frames cycling, %d frames done"
                                 % (args.cpu_seconds, int(r[1]))}
        except Exception as e:  # noqa: BLE001
            cpu = {"value": None, "unit": "Mbit/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (e,)}

    # ---- the same workload through the fixed-point (int16, bit-exact vs AFF3CT's int16 decoder) kernel, device-resident only:
    # a separate process after this one released its HBM (N = 1 runs only; the headline stays the float path the dvbs2
    # binaries run)
    tp = lap("cpu_baseline", tp)
    int16, configs = None, None
    if rank == 0 and world == 1 and args.prec == "f32" and (not args.no_int16 or not args.no_configs):
        if h is not None:
            h.close()
            h = None
            del V, chain, tx
        del batches, sig, rx  # sig / rx: the loop variables still hold the last input batch
        torch.cuda.empty_cache()

        def sub(workload, prec, extra):
            r = subprocess.run([sys.executable, os.path.abspath(__file__), "--prec", prec, "--workload", workload,
                                "--ite", str(args.ite), "--steps", str(max(3, args.steps // 2)), "--warmup", "3", "--no-cpu",
                                "--no-int16", "--no-configs"] + extra, capture_output=True, text=True, timeout=900)
            return json.loads(r.stdout.strip().splitlines()[-1])

        def brief(d):
            return {"value": d["value"], "unit": d["unit"], "dtype": d["dtype"], "workload": d["config"]["workload"],
                    "ms_per_step": d["ms_per_step"], "frames_per_step_per_gpu": d["config"]["frames_per_step_per_gpu"],
                    "stage_ms": d["stage_ms"], "mean_iterations": d["fer"]["mean_iterations"], "ldpc_fer": d["fer"]["ldpc_fer"],
                    "bch_fer": d["fer"]["bch_fer"], "monitor_fer": d["fer"]["monitor"]["fer"], "gpu_launches": d["gpu_launches"],
                    "roofline": {k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_ms", "kernel_share_of_step",
                                                               "algorithmic_bytes_per_frame", "message_passing")}}

        if not args.no_int16:
            try:  # the headline workload through the packed int16 kernel + the int8-LLR socket end to end
                d = sub(args.workload, "i16", (["--ebn0", str(args.ebn0)] if args.ebn0 is not None else []) +
                        ["--e2e-sockets", "llr8_packed", "--e2e-clones", str(args.e2e_clones)])  # 0 = automatic there too
                int16 = brief(d)
                int16["quantizer"] = "Quantizer_pow2 6 bits, 2 fractional (AFF3CT LDPC defaults)"
                int16["message_passing"] = d["roofline"]["message_passing"]
                int16["e2e"] = d.get("e2e")
            except Exception as e:  # noqa: BLE001
                int16 = {"value": None, "error": repr(e)[:300]}
            tp = lap("int16_line", tp)
        if not args.no_configs and args.workload == "qpsk89":
            # BASELINE.json configs 2-4, device-resident: QPSK-S_3/5 at the waterfall (all iterations active), 8PSK-S_3/5 float and
            # fixed point, 8PSK-S_8/9 at high Eb/N0 (divergent iteration counts)
            configs = {}
            for name, wl_, prec in (("qpsk35_f32", "qpsk35", "f32"), ("qpsk35_i16", "qpsk35", "i16"), ("8psk35_f32", "8psk35", "f32"),
                                    ("8psk35_i16", "8psk35", "i16"), ("8psk89_f32", "8psk89", "f32")):
                try:
                    configs[name] = brief(sub(wl_, prec, ["--no-e2e"]))
                except Exception as e:  # noqa: BLE001
                    configs[name] = {"value": None, "error": repr(e)[:300]}
            tp = lap("configs", tp)

    if rank == 0:
        phases["total"] = round(time.time() - T0, 1)
        out = {
            "metric": "info Mbit/s, %s NMS %d-iter decode chain (demap+LDPC+BCH)" % (modcod, args.ite),
            "value": round(value, 2), "unit": "Mbit/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_max / args.steps, 4), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.prec, "data": "synthetic (random payloads through the on-device DVB-S2 TX chain + AWGN, Philox)",
            "config": {"workload": "%s Eb/N0 %.2f dB NMS norm 1.0 %d it, early stop" % (modcod, ebn0, args.ite),
                       "frames_per_step_per_gpu": F, "sigma": round(synth.ebn0_to_sigma(modcod, ebn0), 5),
                       "l2": "inputs %.1f GB per step >> 126 MB L2, %d batches rotated" % (F * 2 * n_sym * 4 / 1e9, NB),
                       "parallelism": "frames sharded by batch, %d rank(s), no data-path collective" % world},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
            "stage_ms": stage_ms, "int16": int16, "configs": configs, "wall_s": phases, "iter_hist": st["iter_hist"][:args.ite + 1],
            "fer": {"frames": int(cnt[0]), "ldpc_fer": cnt[2] / max(1, cnt[0]), "bch_fer": cnt[3] / max(1, cnt[0]),
                    "mean_iterations": cnt[1] / max(1, cnt[0]),
                    "monitor": {"frames": int(mon[0]), "bit_errors": int(mon[1]), "frame_errors": int(mon[2]),
                                "fer": mon[2] / max(1, mon[0]), "ber": mon[1] / max(1, mon[0] * K_bch)}},
        }
        sys.stdout.flush()
        if real_stdout is not None:
            os.dup2(real_stdout, 1)
        print(json.dumps(out), flush=True)
    if h is not None:
        h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
