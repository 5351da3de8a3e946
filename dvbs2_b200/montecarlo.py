"""Device-resident Monte-Carlo BER/FER simulation: the SNR loop of src/mains/TX_RX_BB/main.cpp:139-167 with every task on
the GPU(s) — on-device TX + AWGN (Source .. Channel, main.cpp:75-84), the fused decode chain, BB descrambler and the BFER
monitor — and the stop criterion reduced over ranks.

    python -m dvbs2_b200.montecarlo --mod-cod QPSK-S_8/9 --dec-implem NMS --dec-ite 10 -m 3.8 -M 4.3 -s 0.1 -e 100 -F 16384
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 -m dvbs2_b200.montecarlo ... (one rank per GPU, NCCL)

Flags follow factory::DVBS2 (src/common/Factory/DVBS2/DVBS2.cpp:65-105): -m/-M/-s Eb/N0 range (the loop runs `for ebn0 =
min; ebn0 < max; ebn0 += step`, main.cpp:139), -e max frame errors per point, -F frames per batch (per rank).  A point ends when
the GLOBAL frame-error count reaches -e (Monitor_BFER::fe_limit_achieved, lib/aff3ct/src/Module/Monitor/BFER/Monitor_BFER.cpp:
190-200, reduced like tools::Monitor_reduction_MPI does: lib/aff3ct/include/Tools/Monitor/Monitor_reduction_MPI.hxx:93) or
--max-frames were simulated.  Ranks decode disjoint frames (seeds differ per rank and batch); the only collective is the SUM
of {frames, bit errors, frame errors} after every batch.
"""
import sys
import time

from .dist import Monitor_BFER_reduction, reduce_max


def run_point(step_fn, max_fe, max_frames=0, device=None, group=None, max_batches=1 << 30):
    """One SNR point.  step_fn(batch_index) simulates one batch on this rank and returns the rank's CUMULATIVE counters
    {"n_frames", "n_be", "n_fe"}.  Every rank runs the same number of batches: the reduced counters decide when to stop."""
    red = Monitor_BFER_reduction(max_fe=max_fe, max_frames=max_frames, device=device, group=group)
    b = 0
    while b < max_batches:
        red.update(step_fn(b))
        b += 1
        if red.is_done():
            break
    return red, b


def main(argv=None):
    import argparse
    import os

    import torch
    import torch.distributed as dist

    from . import Decoder_chain_DVBS2_cuda, Handle, Transmitter_DVBS2_cuda, synth
    from .factory import MODCODS

    p = argparse.ArgumentParser(prog="dvbs2_b200.montecarlo")
    p.add_argument("--mod-cod", default="QPSK-S_8/9", choices=MODCODS)
    p.add_argument("--dec-implem", default="NMS", choices=["MS", "NMS"])
    p.add_argument("--dec-ite", type=int, default=50)
    p.add_argument("--dec-norm", type=float, default=1.0)
    p.add_argument("--dec-prec", type=int, default=32, choices=[32, 16, 8])
    p.add_argument("-m", "--sim-noise-min", type=float, default=3.2)
    p.add_argument("-M", "--sim-noise-max", type=float, default=6.0)
    p.add_argument("-s", "--sim-noise-step", type=float, default=0.1)
    p.add_argument("-e", "--max-fe", type=int, default=100)
    p.add_argument("--max-frames", type=int, default=0)
    p.add_argument("-F", "--sim-inter-fra", dest="n_frames", type=int, default=16384)
    p.add_argument("--seed", type=int, default=0xD5B52)
    a = p.parse_args(argv)

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("dvbs2_b200.montecarlo needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    F = a.n_frames
    mc = synth.MODCODS[a.mod_cod]
    h = Handle(modcod=a.mod_cod, n_ite=a.dec_ite, norm=1.0 if a.dec_implem == "MS" else a.dec_norm,
               precision={32: "f32", 16: "i16", 8: "i8"}[a.dec_prec], n_frames=F, device=local)
    tx, chain = Transmitter_DVBS2_cuda(h, F), Decoder_chain_DVBS2_cuda(h, F)
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    sym = torch.zeros((F, 2 * synth.N_LDPC // mc["bps"]), dtype=torch.float32, device=dev)
    V = torch.zeros((F, mc["K_bch"]), dtype=torch.int32, device=dev)
    c1, c2 = torch.zeros(F, dtype=torch.int8, device=dev), torch.zeros(F, dtype=torch.int8, device=dev)
    torch.cuda.synchronize()
    if rank == 0:
        print("# dvbs2_b200 Monte-Carlo: %s, %s %d it, -F %d x %d rank(s), -e %d" % (a.mod_cod, a.dec_implem, a.dec_ite, F, world, a.max_fe))
        print("#   Es/N0 |   Eb/N0 ||        FRA |       BE |       FE |      BER |      FER ||  SIM_THR (Mb/s) | ET (s)")
    ebn0, point = a.sim_noise_min, 0
    while ebn0 < a.sim_noise_max:
        sigma = synth.ebn0_to_sigma(a.mod_cod, ebn0)
        h.reset_stats()  # clears the monitor counters too

        def step(b, sigma=sigma, point=point):
            seed = a.seed + 1000003 * point + 7919 * (b * world + rank)
            tx.random_payload(seed)
            tx.modulate_awgn_dev(sigma, seed + 1, cp, sym)
            chain.demodulate_decode_dev(cp, sym, c1, c2, V)
            tx.check_errors()
            return tx.monitor()  # synchronises this rank's stream

        t0 = time.perf_counter()
        red, n_batches = run_point(step, a.max_fe, a.max_frames, device=dev)
        dt = reduce_max(time.perf_counter() - t0, device=dev)
        if rank == 0:
            t, r = red.total, red.rates(mc["K_bch"])
            esn0 = ebn0 + 10.0 * __import__("math").log10(mc["K_bch"] / synth.N_LDPC * mc["bps"])
            print("  %7.2f | %7.2f || %10d | %8d | %8d | %8.2e | %8.2e || %15.1f | %6.1f" % (
                esn0, ebn0, t["n_frames"], t["n_be"], t["n_fe"], r["ber"], r["fer"], mc["K_bch"] * t["n_frames"] / dt / 1e6, dt))
            sys.stdout.flush()
        ebn0 += a.sim_noise_step
        point += 1
    h.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
