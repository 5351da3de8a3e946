"""FER of the GPU chain against the reference's own Monte-Carlo anchor points (`dvbs2_tx_rx_bb --dec-implem NMS
--dec-ite 10 --dec-simd INTER --ter-freq 0 -e 100`, i.e. norm 1.0, 100 frame errors per point).  The decoders are bit-exact on
identical LLRs (the other GPU suites); this checks the whole Monte-Carlo loop on the device — TX, Philox AWGN, chain, BB
descrambler, BFER monitor — lands inside the joint binomial confidence interval of the reference's points.  QPSK with min-sum
(norm 1) is invariant to the LLR scale, so the reference's estimated sigma versus the true sigma does not matter there; the 8PSK
points use the on-device Estimator_DVBS2 like the reference run; the 16APSK point was taken with the perfect estimator."""
import math
import os

import pytest

from oracle.oracle import MODCODS, N_LDPC, ebn0_to_sigma

pytestmark = pytest.mark.gpu

# (modcod, Eb/N0 dB, reference FER, frames to run here, sigma from the on-device Estimator_DVBS2 as in the reference run)
ANCHORS = [("QPSK-S_8/9", 3.9, 0.129, 8192, False), ("QPSK-S_8/9", 4.0, 1.52e-2, 49152, True), ("QPSK-S_3/5", 2.2, 0.59, 4096, False),
           ("QPSK-S_3/5", 2.5, 2.75e-3, 131072, False), ("8PSK-S_3/5", 3.8, 0.14, 8192, True),
           ("8PSK-S_3/5", 4.0, 1.36e-3, 131072, True), ("8PSK-S_8/9", 6.6, 0.22, 4096, True),
           ("16APSK-S_8/9", 7.6, 8.4e-2, 16384, False)]


@pytest.mark.parametrize("modcod,ebn0,ref_fer,frames,estimate", ANCHORS)
def test_fer_inside_the_reference_confidence_interval(modcod, ebn0, ref_fer, frames, estimate):
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Estimator_DVBS2_cuda, Handle, Transmitter_DVBS2_cuda

    mc = MODCODS[modcod]
    F = 16384 if frames >= 16384 else frames
    h = Handle(modcod=modcod, n_ite=10, n_frames=F)
    tx, ch, est = Transmitter_DVBS2_cuda(h, F), Decoder_chain_DVBS2_cuda(h, F), Estimator_DVBS2_cuda(h, F)
    dev = torch.device("cuda:0")
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    sym = torch.zeros((F, 2 * N_LDPC // mc["bps"]), dtype=torch.float32, device=dev)
    V = torch.zeros((F, mc["K_bch"]), dtype=torch.int32, device=dev)
    c1, c2 = torch.zeros(F, dtype=torch.int8, device=dev), torch.zeros(F, dtype=torch.int8, device=dev)
    sigma = ebn0_to_sigma(modcod, ebn0)
    for b in range(frames // F):
        tx.random_payload(1000 + b)
        tx.modulate_awgn_dev(sigma, 77 + b, cp, sym)
        if estimate:  # the 8PSK points of the reference were taken with --est-type DVBS2 (its default): sigma from M2/M4
            est.estimate_dev(sym, cp)
        ch.demodulate_decode_dev(cp, sym, c1, c2, V, None)
        tx.check_errors()
    mon = tx.monitor()
    h.close()
    assert mon["n_frames"] == frames
    fer = mon["n_fe"] / frames
    # the reference point rests on 100 frame errors (relative sigma 10 %); ours on n_fe; 3.5 sigma of the difference
    rel = math.sqrt(1.0 / 100 + 1.0 / max(1, mon["n_fe"]))
    assert abs(fer - ref_fer) <= 3.5 * rel * ref_fer, (fer, ref_fer, mon)


def test_device_monte_carlo_driver_stops_on_the_frame_error_count():
    """python -m dvbs2_b200.montecarlo: the SNR loop of dvbs2_tx_rx_bb on the device; -e frame errors end a point and the FER
    lands on the reference binary's anchor (QPSK-S_8/9 NMS-10 at 4.0 dB: 1.52e-2 over 13 472 frames)."""
    import subprocess
    import sys

    r = subprocess.run([sys.executable, "-m", "dvbs2_b200.montecarlo", "--mod-cod", "QPSK-S_8/9", "--dec-implem", "NMS",
                        "--dec-ite", "10", "-m", "4.0", "-M", "4.05", "-e", "600", "-F", "8192"], capture_output=True, text=True,
                       timeout=600, cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0, r.stdout + r.stderr
    row = [ln for ln in r.stdout.splitlines() if ln.strip() and not ln.startswith("#")][-1]
    f = [x.strip() for x in row.replace("||", "|").split("|")]
    fra, fe, fer = int(f[2]), int(f[4]), float(f[6])
    assert fe >= 600 and fra % 8192 == 0 and fe - 600 < 8192 * 0.05  # stopped at the first batch that crossed -e
    assert abs(fer - 1.52e-2) < 4 * (1.52e-2 * (1 - 1.52e-2) * (1 / fra + 1 / 13472)) ** 0.5
