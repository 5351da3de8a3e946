#!/usr/bin/env python
"""Generates tests/golden/rx_front.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libdvbs2_ref.so): per MODCOD two PL
frames made by the reference's Framer::generate from noisy symbols, the sha256 of Framer::remove_plh's output on them and the
three outputs of Estimator_DVBS2::estimate on that output.  Run in the build container:
    python tests/golden/make_golden_rx_front.py
Pins oracle_remove_plh / oracle_estimate (and, through tests/test_gpu_golden.py, the CUDA kernels) where the reference tree is
absent."""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import MODCODS, N_LDPC, Oracle, Ref, synth_frames  # noqa: E402

EBN0 = {"QPSK-S_8/9": 4.0, "QPSK-S_3/5": 2.2, "8PSK-S_3/5": 3.9, "8PSK-S_8/9": 7.5, "16APSK-S_8/9": 7.6}


def main():
    r, o = Ref(), Oracle()
    out = {}
    for name, mc in MODCODS.items():
        n_sym = N_LDPC // mc["bps"]
        pl_sym = 90 * (n_sym // 90 + 1) + 36 * (n_sym // 1440)  # DVBS2.cpp:352-355
        s = synth_frames(o, name, 2, EBN0[name], seed=77)       # the oracle TX is itself pinned to the reference (chain_*.npz)
        pl = r.framer_generate(name, n_sym, pl_sym, s["rx"])
        x = r.remove_plh(name, n_sym, pl_sym, pl)
        sig, eb, es = r.estimate(n_sym, mc["K_bch"], mc["bps"], x)
        k = name.replace("/", "_")
        out[k + "_pl"] = pl.astype(np.float32)
        out[k + "_xfec_sha"] = np.frombuffer(hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest(), np.uint8)
        out[k + "_est"] = np.stack([sig, eb, es]).astype(np.float32)
    np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "rx_front.npz"), **out)
    print("wrote rx_front.npz", {k: v.shape for k, v in out.items() if k.endswith("_pl")})


if __name__ == "__main__":
    main()
