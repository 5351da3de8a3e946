"""Reproducible LDPC-stage inputs for the 256-frame golden fixtures (tests/golden/ldpc256_*.npz).

Everything here is integer hashing plus IEEE-754 float32 add / multiply / min done element-wise by numpy in a fixed order — no
libm, no platform-dependent reduction — so the LLRs can be REGENERATED bit for bit on any host instead of being stored (a
256-frame float32 LLR set would be 16.6 MB per MODCOD).  The fixture keeps their SHA-256 so that drift is detected.

    payload bits   splitmix64(seed, frame, i) & 1
    TX             oracle.tx (pinned bit for bit to the reference's TX chain by test_oracle_vs_reference / tables.npz)
    channel        y = c + sigma * n,  n = (sum of 12 hashed uniforms) - 6   (CLT Gaussian, |n| <= 6)
    LLR            max-log demapper over the MODCOD's constellation, then the column-row deinterleaver
(The reference's log-MAP demapper is covered by chain_*.npz; what matters here is a large, realistic and exactly
reproducible set of decoder inputs across the waterfall.)
"""
import hashlib

import numpy as np

N_LDPC = 16200
# Eb/N0 points (dB) and frames per point: FER from ~1 down to ~1e-2 .. 0 per MODCOD, 256 frames in all
POINTS = {
    "QPSK-S_8/9": ((3.7, 3.95, 4.25), (96, 96, 64)),
    "QPSK-S_3/5": ((2.0, 2.3, 2.65), (96, 96, 64)),
    "8PSK-S_3/5": ((3.5, 3.85, 4.15), (96, 96, 64)),
    "8PSK-S_8/9": ((6.4, 6.7, 7.1), (96, 96, 64)),
    "16APSK-S_8/9": ((7.4, 7.75, 8.1), (96, 96, 64)),
}


def _mix(x):
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)).astype(np.uint64)
    x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)).astype(np.uint64)
    return x ^ (x >> np.uint64(31))


def hashed(seed, frame, idx, k=0):
    """splitmix64 of (seed, frame, idx, k) -> uint64 array shaped like the broadcast of frame and idx"""
    with np.errstate(over="ignore"):
        x = _mix(np.uint64(seed) + np.uint64(0x632BE59BD9B4E019) * np.uint64(k))
        x = _mix(x ^ frame.astype(np.uint64))
        return _mix(x ^ (idx.astype(np.uint64) << np.uint64(20)))


def payload(seed, n_frames, K):
    f = np.arange(n_frames, dtype=np.uint64)[:, None]
    i = np.arange(K, dtype=np.uint64)[None, :]
    return (hashed(seed, f, i) & np.uint64(1)).astype(np.int32)


def gauss(seed, n_frames, n):
    """float32 [n_frames][n]: sum of 12 uniforms in (0,1) minus 6, added in a fixed order"""
    f = np.arange(n_frames, dtype=np.uint64)[:, None]
    i = np.arange(n, dtype=np.uint64)[None, :]
    acc = np.zeros((n_frames, n), np.float32)
    for k in range(1, 13):
        u = ((hashed(seed, f, i, k) >> np.uint64(41)).astype(np.float32) + np.float32(0.5)) * np.float32(2.0 ** -23)
        acc = acc + u
    return acc - np.float32(6.0)


def maxlog_llr(cstl, sigma, rx, bps):
    """rx [F][2*n_sym] float32, sigma [F] -> LLR [F][16200] in transmitted (interleaved) bit order, positive = bit 0"""
    yr, yi = rx[:, 0::2], rx[:, 1::2]
    inv = (np.float32(1.0) / (np.float32(2.0) * sigma * sigma)).astype(np.float32)[:, None]
    big = np.float32(3.0e38)
    m0 = [np.full(yr.shape, big, np.float32) for _ in range(bps)]
    m1 = [np.full(yr.shape, big, np.float32) for _ in range(bps)]
    for j in range(1 << bps):
        dr, di = yr - np.float32(cstl[j, 0]), yi - np.float32(cstl[j, 1])
        d = dr * dr + di * di
        for b in range(bps):
            if (j >> b) & 1:
                m1[b] = np.minimum(m1[b], d)
            else:
                m0[b] = np.minimum(m0[b], d)
    llr = np.empty((rx.shape[0], N_LDPC), np.float32)
    for b in range(bps):
        llr[:, b::bps] = (m1[b] - m0[b]) * inv
    return llr


def make(oracle, modcod, seed=0x5EED, points=None):
    """-> dict(U, ldpc (codewords), llr (natural order, float32), ebn0 per frame)"""
    from oracle.oracle import MODCODS, ebn0_to_sigma

    mc = MODCODS[modcod]
    ebn0s, counts = points or POINTS[modcod]
    F = int(sum(counts))
    U = payload(seed, F, mc["K_bch"])
    tx = oracle.tx(modcod, U)
    sig = np.concatenate([np.full(c, ebn0_to_sigma(modcod, e), np.float32) for e, c in zip(ebn0s, counts)])
    n = gauss(seed + 1, F, tx["sym"].shape[1])
    rx = tx["sym"] + sig[:, None] * n
    llr = maxlog_llr(oracle.constellation(mc["bps"]), sig, rx.astype(np.float32), mc["bps"])
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    nat = np.ascontiguousarray(llr[:, inv])
    return dict(U=U, bch=tx["bch"], ldpc=tx["ldpc"], llr=nat, sigma=sig,
                ebn0=np.concatenate([np.full(c, e, np.float32) for e, c in zip(ebn0s, counts)]))


def quantize(llr, dec, bits):
    """Quantizer_pow2 (Quantizer_pow2.cpp:184-190): sat(round-half-away(2^dec * y), +-(2^(bits-1) - 1)) as int16"""
    y = llr * np.float32(1 << dec)
    r = np.sign(y) * np.floor(np.abs(y) + np.float32(0.5))
    lim = (1 << (bits - 1)) - 1
    return np.clip(r, -lim, lim).astype(np.int16)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def load_fixture(oracle, modcod):
    """tests/golden/ldpc256_<modcod>.npz (reference outputs, make_golden_ldpc.py) next to its regenerated inputs:
    -> (npz, inputs dict, q16, q8, dict of expected bit arrays per decoder key)"""
    import os

    from oracle.oracle import MODCODS

    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "ldpc256_%s.npz" % modcod.replace("/", "_")))
    d = make(oracle, modcod, seed=int(g["seed"]))
    q16, q8 = quantize(d["llr"], 2, 6), quantize(d["llr"], 1, 6).astype(np.int8)
    # the inputs are rebuilt from integer hashing + IEEE float32 add/mul/min: a mismatch means numpy changed, not the decoder
    assert sha(d["llr"]) == str(g["llr_sha"]) and sha(q16) == str(g["q16_sha"]) and sha(q8) == str(g["q8_sha"])
    N, K = MODCODS[modcod]["N_bch"], MODCODS[modcod]["K_bch"]
    ub = lambda a, n: np.unpackbits(a, axis=-1)[..., :n].astype(np.int32)  # noqa: E731
    want = {k[:-2]: ub(g[k], N) ^ d["ldpc"][:, :N] for k in g.files if k.endswith("_x") and k != "bch_x"}
    want["bch"] = ub(g["bch_x"], K) ^ d["bch"][:, :K]
    return g, d, q16, q8, want
