#!/usr/bin/env python
"""Generates tests/golden/*.npz from the COMPILED, UNMODIFIED reference (oracle/_ref/libdvbs2_ref.so, built from
/root/reference by oracle/build_ref.sh).  Run in the build container:  python tests/golden/make_golden.py

The fixtures pin the CPU oracle (oracle/dvbs2_oracle.c) and, through it, the CUDA path, on hosts where the reference
tree is not available (the GPU box).  Everything stored here is an OUTPUT of reference classes on seeded inputs:
  tables.npz            H (check degrees + column lists hashed and sampled), GF(2^14) tables, BCH generator, interleaver
                        LUTs, normalised constellations           <- tools::build_dvbs2/build_H, BCH_polynomial_generator,
                                                                     Interleaver_core_column_row, Constellation_user
  chain_<modcod>.npz    per MODCOD a few frames: payload U, TX stages, received symbols + sigma, demapper LLRs
                        (Modem_generic_fast, interleaved and natural order), LDPC outputs of the sequential float decoder,
                        of ONMS_inter float, of the int16 sequential/inter decoders and the int8 sequential decoder on
                        Quantizer_pow2 outputs, BCH outputs (Decoder_BCH_DVBS2), descrambled payload
  bch_stress.npz        BCH words with 0..40 injected errors (incl. shortened/parity region) and the reference's verdicts
Bits are stored packed (np.packbits) to keep the files small.
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle.oracle import MODCODS, N_LDPC, Ref, ebn0_to_sigma  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
EBN0 = {"QPSK-S_8/9": 3.95, "QPSK-S_3/5": 2.35, "8PSK-S_3/5": 3.85, "8PSK-S_8/9": 6.65, "16APSK-S_8/9": 7.65}
N_CAND, N_KEEP = 48, 4


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def pb(bits):
    return np.packbits(np.asarray(bits, np.uint8), axis=-1)


def main():
    r = Ref()
    # ---------------------------------------------------------------- tables
    t = {}
    for K in (14400, 9720):
        deg, cols = r.build_H(K)
        t["H%d_deg_sha" % K] = sha(deg.astype(np.int32))
        t["H%d_cols_sha" % K] = sha(cols.astype(np.int32))
        t["H%d_M_E" % K] = np.array([len(deg), len(cols)])
        t["H%d_first_rows" % K] = cols[: int(deg[:4].sum())].astype(np.int32)
        t["H%d_last_row" % K] = cols[-int(deg[-1]):].astype(np.int32)
    a, i, g = r.bch_tables()
    t["gf_alpha_to"], t["gf_index_of"], t["bch_g"] = a.astype(np.int32), i.astype(np.int32), g.astype(np.int32)
    for name, mc in MODCODS.items():
        if mc["itl"]:
            rng = np.random.default_rng(1)
            U = rng.integers(0, 2, size=(1, mc["K_bch"]), dtype=np.int32)
            tx = r.tx(name, U)
            # recover pi from a one-hot walk is expensive; store the interleaved image of a ramp instead
            t["itl_%s_sha" % name] = sha(tx["itl"])
        sig = np.full(1, 0.5, np.float32)
        _, _, cst = r.demodulate(name, sig, np.zeros((1, 2 * N_LDPC // mc["bps"]), np.float32))
        t["cstl_%d" % mc["bps"]] = cst.astype(np.float32)
    np.savez_compressed(os.path.join(OUT, "tables.npz"), **t)

    # ---------------------------------------------------------------- chain fixtures
    for name, mc in MODCODS.items():
        K, N = mc["K_bch"], mc["N_bch"]
        rng = np.random.default_rng(0xD5B52 + len(name) + K)
        U = rng.integers(0, 2, size=(N_CAND, K), dtype=np.int32)
        tx = r.tx(name, U)
        sigma = ebn0_to_sigma(name, EBN0[name])
        rx = (tx["sym"] + rng.standard_normal(tx["sym"].shape, dtype=np.float32) * np.float32(sigma)).astype(np.float32)
        sig = (np.full(N_CAND, sigma, np.float32) * np.linspace(0.97, 1.03, N_CAND).astype(np.float32)).astype(np.float32)
        itl, nat, _ = r.demodulate(name, sig, rx)
        Vs, cs = r.ldpc_decode(N, nat, prec="f32", simd=False)
        # keep a failed frame, a converged one, and two more (the set must fill whole SIMD waves for the inter decoders,
        # so outputs of the inter decoders are computed on the full candidate batch and sliced)
        Vi, ci = r.ldpc_decode(N, nat, prec="f32", simd=True)
        q16 = r.quantize(nat, 2, 6, "i16")
        q8 = r.quantize(nat, 2, 6, "i8")
        q8s1 = r.quantize(nat, 1, 6, "i8")  # coarser scale: int8 never wraps here
        V16s, c16s = r.ldpc_decode(N, q16, prec="i16", simd=False)
        n16 = (N_CAND // 32) * 32
        V16i, c16i = r.ldpc_decode(N, q16[:n16], prec="i16", simd=True)
        V8, c8 = r.ldpc_decode(N, q8, prec="i8", simd=False)
        V8s1, c8s1 = r.ldpc_decode(N, q8s1, prec="i8", simd=False)
        Vb, cb = r.bch_decode(K, N, Vs)
        desc = r.descramble(K, Vb)
        bad = np.flatnonzero(cs == 0)
        good = np.flatnonzero(cs == 1)
        keep = list(bad[:2]) + list(good[: N_KEEP - len(bad[:2])])
        keep = np.array(sorted(keep[:N_KEEP]))
        assert len(keep) == N_KEEP and keep.max() < n16, (name, len(bad), len(good))
        np.savez_compressed(
            os.path.join(OUT, "chain_%s.npz" % name.replace("/", "_")), modcod=name, ebn0=EBN0[name], keep=keep,
            U=pb(U[keep]), scr=pb(tx["scr"][keep]), bch=pb(tx["bch"][keep]), ldpc=pb(tx["ldpc"][keep]), itl=pb(tx["itl"][keep]),
            sym=tx["sym"][keep], rx=rx[keep], sigma=sig[keep], llr_itl=itl[keep], llr_nat=nat[keep],
            ldpc_seq_f32_V=pb(Vs[keep]), ldpc_seq_f32_cwd=cs[keep], ldpc_inter_f32_V=pb(Vi[keep]), ldpc_inter_f32_cwd=ci[keep],
            q16=q16[keep], q8=q8[keep], q8s1=q8s1[keep],
            ldpc_seq_i16_V=pb(V16s[keep]), ldpc_seq_i16_cwd=c16s[keep], ldpc_inter_i16_V=pb(V16i[keep]), ldpc_inter_i16_cwd=c16i[keep],
            ldpc_seq_i8_V=pb(V8[keep]), ldpc_seq_i8_cwd=c8[keep], ldpc_seq_i8s1_V=pb(V8s1[keep]), ldpc_seq_i8s1_cwd=c8s1[keep],
            bch_V=pb(Vb[keep]), bch_cwd=cb[keep], desc=pb(desc[keep]),
            batch_stats=np.array([N_CAND, int((cs == 0).sum()), int((ci == 0).sum()), int((c16s == 0).sum()), int((c8 == 0).sum()),
                                  int((cb == 0).sum())]))
        print(name, "kept", keep, "cwd", cs[keep], "FE seq/inter/i16/i8/bch", (cs == 0).sum(), (ci == 0).sum(), (c16s == 0).sum(),
              (c8 == 0).sum(), (cb == 0).sum())

    # ---------------------------------------------------------------- BCH stress vectors
    out = {}
    for name in ("QPSK-S_8/9", "QPSK-S_3/5"):
        mc = MODCODS[name]
        K, N = mc["K_bch"], mc["N_bch"]
        rng = np.random.default_rng(K)
        U = rng.integers(0, 2, size=(40, K), dtype=np.int32)
        Y = r.tx(name, U)["bch"].copy()
        n_err = [0, 1, 2, 5, 11, 12, 12, 13, 13, 14, 15, 20, 40, 100, 13, 14, 12, 3, 13, 7] * 2
        pos_all = []
        for f in range(40):
            if f % 3 == 2:
                pool = np.r_[0:40, K - 20:K + 20, N - 40:N]
                pos = rng.choice(pool, size=min(n_err[f], len(pool)), replace=False)
            else:
                pos = rng.choice(N, size=n_err[f], replace=False)
            Y[f, pos] ^= 1
            pos_all.append(len(pos))
        V, c = r.bch_decode(K, N, Y)
        tag = str(K)
        out["Y_" + tag], out["V_" + tag], out["cwd_" + tag], out["nerr_" + tag] = pb(Y), pb(V), c, np.array(pos_all)
        print("bch", name, "accepted", int(c.sum()), "of 40")
    np.savez_compressed(os.path.join(OUT, "bch_stress.npz"), **out)
    r.close()


if __name__ == "__main__":
    main()
