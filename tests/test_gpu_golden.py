"""GPU parity against the committed golden vectors of the compiled reference (tests/golden, made by make_golden.py):
the CUDA path through the C ABI must reproduce the reference's outputs on the reference's own inputs."""
import os

import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


def ub(packed, n):
    return np.unpackbits(packed, axis=-1)[..., :n].astype(np.int32)


def load(modcod):
    return np.load(os.path.join(GOLD, "chain_%s.npz" % modcod.replace("/", "_")))


def _ldpc(modcod, Y, prec, **kw):
    from dvbs2_b200 import Decoder_LDPC_DVBS2_cuda, Handle

    F = Y.shape[0]
    h = Handle(modcod=modcod, precision=prec, n_frames=F, **kw)
    dec = Decoder_LDPC_DVBS2_cuda(h, F)
    V, cwd = np.zeros((F, dec.K), np.int32), np.zeros(F, np.int8)
    dec.decode_siho(np.ascontiguousarray(Y), cwd, V)
    h.close()
    return V, cwd


@pytest.mark.parametrize("modcod", ALL)
def test_ldpc_all_precisions_vs_reference_outputs(modcod):
    g, N = load(modcod), MODCODS[modcod]["N_bch"]
    for nf, coop in ((1, 0), (1, -1), (2, 0)):  # cooperative small-batch kernel, 32- and 64-frame groups of the lane = frame kernel
        V, cwd = _ldpc(modcod, g["llr_nat"], "f32", ldpc_frames_per_lane=nf, ldpc_coop_max_frames=coop)
        assert np.array_equal(cwd, g["ldpc_seq_f32_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_f32_V"], N))
        ok = g["ldpc_inter_f32_cwd"] == 1  # float NMS of --dec-simd INTER: identical on every frame that converges
        assert np.array_equal(V[ok], ub(g["ldpc_inter_f32_V"], N)[ok])
    V, cwd = _ldpc(modcod, g["q16"], "i16")
    assert np.array_equal(cwd, g["ldpc_seq_i16_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_i16_V"], N))
    ok = g["ldpc_inter_i16_cwd"] == 1
    assert np.array_equal(V[ok], ub(g["ldpc_inter_i16_V"], N)[ok])
    V, cwd = _ldpc(modcod, g["q8"], "i8")
    assert np.array_equal(cwd, g["ldpc_seq_i8_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_i8_V"], N))
    V, cwd = _ldpc(modcod, g["q8s1"], "i8", quant_dec=1)
    assert np.array_equal(cwd, g["ldpc_seq_i8s1_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_i8s1_V"], N))


@pytest.mark.parametrize("modcod", ALL)
def test_demapper_quantiser_bch_tx_vs_reference_outputs(modcod):
    import ctypes as C

    import torch

    from dvbs2_b200 import Decoder_BCH_DVBS2_cuda, Handle, Interleaver_DVBS2_cuda, Modem_DVBS2_cuda, Transmitter_DVBS2_cuda

    g, mc = load(modcod), MODCODS[modcod]
    K, N = mc["K_bch"], mc["N_bch"]
    F = g["rx"].shape[0]
    h = Handle(modcod=modcod, n_frames=F)
    llr = np.zeros((F, N_LDPC), np.float32)
    Modem_DVBS2_cuda(h, F).demodulate(np.ascontiguousarray(g["sigma"]), np.ascontiguousarray(g["rx"]), llr)
    err = np.abs(llr - g["llr_itl"]) / np.maximum(1.0, np.abs(g["llr_itl"]))
    assert err.max() <= 1e-4, err.max()
    nat = np.zeros_like(llr)
    Interleaver_DVBS2_cuda(h, F).deinterleave(np.ascontiguousarray(g["llr_itl"]), nat)
    assert np.array_equal(nat, g["llr_nat"])
    V, cwd = np.zeros((F, K), np.int32), np.zeros(F, np.int8)
    Decoder_BCH_DVBS2_cuda(h, F).decode_hiho(ub(g["ldpc_seq_f32_V"], N), cwd, V)
    assert np.array_equal(cwd, g["bch_cwd"]) and np.array_equal(V, ub(g["bch_V"], K))
    tx = Transmitter_DVBS2_cuda(h, F)
    tx.set_payload(ub(g["U"], K))
    dev = torch.device("cuda:0")
    cp, sym = torch.zeros(F, device=dev), torch.zeros((F, 2 * N_LDPC // mc["bps"]), device=dev)
    tx.modulate_awgn_dev(0.0, 0, cp, sym)
    assert np.array_equal(tx.get_bits("ldpc"), ub(g["ldpc"], N_LDPC)) and np.array_equal(sym.cpu().numpy(), g["sym"])
    h.close()
    for prec, key, dec in (("i16", "q16", 2), ("i8", "q8", 2), ("i8", "q8s1", 1)):
        hq = Handle(modcod=modcod, n_frames=F, precision=prec, quant_dec=dec)
        out = np.zeros(g["llr_nat"].shape, np.int16 if prec == "i16" else np.int8)
        y = np.ascontiguousarray(g["llr_nat"])
        hq.check(hq.lib.dvbs2_cuda_quantize(hq._h, y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), y.size))
        assert np.array_equal(out, g[key])
        hq.close()


def test_bch_stress_vectors_vs_reference_outputs():
    from dvbs2_b200 import Decoder_BCH_DVBS2_cuda, Handle

    g = np.load(os.path.join(GOLD, "bch_stress.npz"))
    for modcod, K, N in (("QPSK-S_8/9", 14232, 14400), ("QPSK-S_3/5", 9552, 9720)):
        Y = np.ascontiguousarray(ub(g["Y_%d" % K], N))
        F = Y.shape[0]
        h = Handle(modcod=modcod, n_frames=F)
        V, cwd = np.zeros((F, K), np.int32), np.zeros(F, np.int8)
        Decoder_BCH_DVBS2_cuda(h, F).decode_hiho(Y, cwd, V)
        assert np.array_equal(cwd, g["cwd_%d" % K]) and np.array_equal(V, ub(g["V_%d" % K], K))
        h.close()


@pytest.mark.parametrize("modcod", ALL)
def test_framer_and_estimator_vs_reference_outputs(modcod):
    """remove_plh and Estimator_DVBS2 on the device against the compiled reference's outputs (rx_front.npz)."""
    import hashlib

    from dvbs2_b200 import Estimator_DVBS2_cuda, Framer_DVBS2_cuda, Handle

    g = np.load(os.path.join(GOLD, "rx_front.npz"))
    k, mc = modcod.replace("/", "_"), MODCODS[modcod]
    n_sym = N_LDPC // mc["bps"]
    pl = np.ascontiguousarray(g[k + "_pl"])
    F = pl.shape[0]
    h = Handle(modcod=modcod, n_frames=F)
    x = np.zeros((F, 2 * n_sym), np.float32)
    Framer_DVBS2_cuda(h, F).remove_plh(pl, x)
    assert hashlib.sha256(x.tobytes()).digest() == g[k + "_xfec_sha"].tobytes()
    sig, eb, es = (np.zeros(F, np.float32) for _ in range(3))
    Estimator_DVBS2_cuda(h, F).estimate(x, sig, eb, es)
    ref = g[k + "_est"]
    assert np.max(np.abs(sig - ref[0]) / ref[0]) < 5e-6
    assert np.allclose(eb, ref[1], rtol=0, atol=2e-4) and np.allclose(es, ref[2], rtol=0, atol=2e-4)
    h.close()


@pytest.mark.parametrize("modcod", ALL)
def test_ldpc_256_frames_per_modcod_vs_reference_outputs(oracle, modcod):
    """tests/golden/ldpc256_*.npz: 256 frames per MODCOD across the waterfall (FER ~1 -> 0), outputs of the compiled reference
    classes.  The CUDA path must give the sequential decoders' bits and CWD on every frame (float through all three kernel
    shapes, int16, int8), the oracle-pinned iteration counts, the bits of both inter-frame SIMD decoders (ONMS_inter a8, the
    templated ..._inter a9) wherever those converge, and the BCH stage's verdicts behind it."""
    import sys

    sys.path.insert(0, GOLD)
    import ldpc_inputs as gi
    from dvbs2_b200 import Decoder_BCH_DVBS2_cuda, Decoder_LDPC_DVBS2_cuda, Handle

    g, d, q16, q8, want = gi.load_fixture(oracle, modcod)
    mc = MODCODS[modcod]
    F, N, K = 256, mc["N_bch"], mc["K_bch"]

    def run(Y, prec, **kw):
        h = Handle(modcod=modcod, precision=prec, n_frames=F, **kw)
        V, cwd, nit = np.zeros((F, N), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int32)
        Decoder_LDPC_DVBS2_cuda(h, F).decode_siho(np.ascontiguousarray(Y), cwd, V, nit)
        h.close()
        return V, cwd, nit

    for nf, coop, onchip in ((1, 0, -1), (1, -1, -1), (2, -1, -1), (0, 0, 1)):  # the last: the on-chip kernel (design B)
        V, cwd, nit = run(d["llr"], "f32", ldpc_frames_per_lane=nf, ldpc_coop_max_frames=coop, ldpc_onchip=onchip)
        assert np.array_equal(cwd, g["seq_f32_cwd"]) and np.array_equal(V, want["seq_f32"]) and np.array_equal(nit, g["seq_f32_nit"])
        for key in ("onms_f32", "tint_f32"):
            ok = g[key + "_cwd"] == 1
            assert np.array_equal(V[ok], want[key][ok])
    V16, c16, _ = run(q16, "i16")
    assert np.array_equal(c16, g["seq_i16_cwd"]) and np.array_equal(V16, want["seq_i16"])
    for key in ("onms_i16", "tint_i16"):
        ok = g[key + "_cwd"] == 1
        assert np.array_equal(V16[ok], want[key][ok])
    V8, c8, _ = run(q8, "i8", quant_dec=1)
    assert np.array_equal(c8, g["seq_i8_cwd"]) and np.array_equal(V8, want["seq_i8"])
    h = Handle(modcod=modcod, n_frames=F)
    Vb, cb = np.zeros((F, K), np.int32), np.zeros(F, np.int8)
    Decoder_BCH_DVBS2_cuda(h, F).decode_hiho(np.ascontiguousarray(want["seq_f32"]), cb, Vb)
    h.close()
    assert np.array_equal(cb, g["bch_cwd"]) and np.array_equal(Vb, want["bch"])
