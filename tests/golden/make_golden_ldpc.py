#!/usr/bin/env python
"""Generates tests/golden/ldpc256_<modcod>.npz: 256 frames per MODCOD across the waterfall, decoded by the COMPILED,
UNMODIFIED reference classes (oracle/_ref/libdvbs2_ref.so):

  seq_f32    Decoder_LDPC_BP_horizontal_layered<int,float,Update_rule_NMS<float>>           
  onms_f32   Decoder_LDPC_BP_horizontal_layered_ONMS_inter<int,float>  (--dec-simd INTER)   (a8)
  tint_f32   Decoder_LDPC_BP_horizontal_layered_inter<int,float,Update_rule_NMS_simd<float>> (a9, C++17 builds)
  seq_i16 / onms_i16 / tint_i16   the same three on Quantizer_pow2(6, 2) outputs, R = int16_t
  seq_i8     sequential int8 decoder on Quantizer_pow2(6, 1) outputs
  bch        Decoder_BCH_DVBS2 on seq_f32's bits

The decoder INPUTS are not stored: tests/golden/ldpc_inputs.py regenerates them bit for bit (their SHA-256 is stored).
Bits are stored as XOR against the transmitted codeword / payload, packed — mostly zeros, so the files stay small.
Run in the build container:  python tests/golden/make_golden_ldpc.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ldpc_inputs as gi  # noqa: E402
from oracle.oracle import MODCODS, Oracle, Ref  # noqa: E402


def pb(bits):
    return np.packbits(np.asarray(bits, np.uint8), axis=-1)


def main():
    o, r = Oracle(), Ref()
    nthr = os.cpu_count() or 4
    for name, mc in MODCODS.items():
        K, N = mc["K_bch"], mc["N_bch"]
        d = gi.make(o, name)
        llr, cw = d["llr"], d["ldpc"][:, :N]
        q16, q8 = gi.quantize(llr, 2, 6), gi.quantize(llr, 1, 6).astype(np.int8)
        assert np.array_equal(q16, r.quantize(llr, 2, 6, "i16")) and np.array_equal(q8, r.quantize(llr, 1, 6, "i8"))
        out = dict(modcod=name, seed=0x5EED, ebn0=d["ebn0"], llr_sha=gi.sha(llr), q16_sha=gi.sha(q16), q8_sha=gi.sha(q8))
        for key, prec, simd, Y in (("seq_f32", "f32", False, llr), ("onms_f32", "f32", True, llr), ("tint_f32", "f32", "tinter", llr),
                                   ("seq_i16", "i16", False, q16), ("onms_i16", "i16", True, q16), ("tint_i16", "i16", "tinter", q16)):
            V, c = r.ldpc_decode(N, Y, prec=prec, simd=simd, threads=nthr)
            out[key + "_x"], out[key + "_cwd"] = pb(V ^ cw), c
        V, c = r.ldpc_decode(N, q8, prec="i8")
        out["seq_i8_x"], out["seq_i8_cwd"] = pb(V ^ cw), c
        Vs = np.unpackbits(out["seq_f32_x"], axis=-1)[:, :N].astype(np.int32) ^ cw
        Vb, cb = r.bch_decode(K, N, Vs, threads=min(4, nthr))
        out["bch_x"], out["bch_cwd"] = pb(Vb ^ d["bch"][:, :K]), cb
        # iteration counts are not observable on the reference classes; the oracle's (bit-exact on V and CWD, checked here)
        Vo, co, no = o.ldpc_decode_mt(N, llr, nthr)
        assert np.array_equal(Vo, Vs) and np.array_equal(co, out["seq_f32_cwd"])
        out["seq_f32_nit"] = no.astype(np.int8)
        np.savez_compressed(os.path.join(HERE, "ldpc256_%s.npz" % name.replace("/", "_")), **out)
        fe = lambda k: [int((out[k][a:b] == 0).sum()) for a, b in ((0, 96), (96, 192), (192, 256))]  # noqa: E731
        print(name, "FE per point seq_f32", fe("seq_f32_cwd"), "onms", fe("onms_f32_cwd"), "tint", fe("tint_f32_cwd"), "i16",
              fe("seq_i16_cwd"), "i8", fe("seq_i8_cwd"), "bch", fe("bch_cwd"), "mean it", float(no.mean()))
    r.close()


if __name__ == "__main__":
    main()
