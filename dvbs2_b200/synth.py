"""Synthetic AWGN frames for benchmarks and Monte-Carlo runs, generated on the device with torch.

The DVB-S2 codes are linear and the chain (max* demapper, min-sum, BCH) is symmetric under codeword translation on a
symmetric channel, so the all-zero codeword through AWGN is the same decoding workload — iteration counts, FER, memory
traffic — as any other codeword (the TX chain of D/mains/TX_RX_BB/main.cpp:75-84 would produce).  Noise model and
Eb/N0 -> sigma follow A/src/Tools/general_utils.cpp:78-89,115-135 as used by main.cpp:142-148:
    EsN0 = EbN0 + 10 log10(K_bch / 16200 * bps),  sigma = sqrt(1 / (2 * 10^(EsN0/10))),  y = c + sigma (n_re + i n_im).
"""
import math

N_LDPC = 16200

# name -> bps, K_bch, N_bch (DVBS2.cpp:287-356) and constellation point 0 (conf/mod/*.mod, before normalisation)
MODCODS = {
    "QPSK-S_8/9": dict(bps=2, K_bch=14232, N_bch=14400),
    "QPSK-S_3/5": dict(bps=2, K_bch=9552, N_bch=9720),
    "8PSK-S_3/5": dict(bps=3, K_bch=9552, N_bch=9720),
    "8PSK-S_8/9": dict(bps=3, K_bch=14232, N_bch=14400),
    "16APSK-S_8/9": dict(bps=4, K_bch=14232, N_bch=14400),
}

_C16 = [(0.797081141330456, 0.797081141330456), (0.291751946604441, 1.08883308793490),
        (1.08883308793490, 0.291751946604441), (0.306569669742483, 0.306569669742483)]


def point0(bps):
    """Constellation point of the all-zero label, normalised to unit mean energy (Constellation_user.hxx:22-67)."""
    if bps == 2:
        return (1.0 / math.sqrt(2.0), 1.0 / math.sqrt(2.0))
    if bps == 3:
        return (0.707106781, 0.707106781)
    if bps == 4:
        e = sum(x * x + y * y for x, y in _C16) / 4.0  # the 16 points are the 4 above in the 4 quadrants
        s = 1.0 / math.sqrt(e)
        return (_C16[0][0] * s, _C16[0][1] * s)
    raise ValueError(bps)


def ebn0_to_sigma(modcod, ebn0_db):
    mc = MODCODS[modcod]
    esn0 = ebn0_db + 10.0 * math.log10(mc["K_bch"] / N_LDPC * mc["bps"])
    return math.sqrt(1.0 / (2.0 * 10.0 ** (esn0 / 10.0)))


def awgn_zero_codeword(modcod, n_frames, ebn0_db, device, seed=0xD5B52, chunk=4096):
    """Received symbols [n_frames][2*n_sym] (interleaved I/Q, float32) and sigma[n_frames] on `device`."""
    import torch

    mc = MODCODS[modcod]
    n_sym = N_LDPC // mc["bps"]
    sigma = ebn0_to_sigma(modcod, ebn0_db)
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    rx = torch.empty((n_frames, 2 * n_sym), dtype=torch.float32, device=device)
    re, im = point0(mc["bps"])
    for f0 in range(0, n_frames, chunk):
        blk = rx[f0:f0 + chunk]
        blk.normal_(0.0, sigma, generator=gen)
        blk[:, 0::2] += re
        blk[:, 1::2] += im
    sig = torch.full((n_frames,), sigma, dtype=torch.float32, device=device)
    return sig, rx
