"""GPU parity: BCH decoder (bit-exact), demapper/deinterleaver/quantiser (float tolerance / exact), fused chain."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, synth_frames

pytestmark = pytest.mark.gpu

ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


def _bch_stress(oracle, K, N, F, seed):
    rng = np.random.default_rng(seed)
    U = rng.integers(0, 2, size=(F, K), dtype=np.int32)
    Y = oracle.bch_encode(K, N, U)
    pattern = [0, 1, 2, 5, 11, 12, 12, 13, 13, 14, 15, 20, 40, 100, 13, 14, 12, 3, 13, 7]
    for f in range(F):
        ne = pattern[f % len(pattern)]
        if f % 3 == 2:  # crowd the errors into the head, the tail and the parity section
            pool = np.r_[0:40, K - 20:K + 20, N - 40:N]
            pos = rng.choice(pool, size=min(ne, len(pool)), replace=False)
        else:
            pos = rng.choice(N, size=ne, replace=False)
        Y[f, pos] ^= 1
    return U, np.ascontiguousarray(Y)


@pytest.mark.parametrize("modcod", ["QPSK-S_8/9", "QPSK-S_3/5"])
@pytest.mark.parametrize("F", [1, 33, 200])
def test_bch_bit_exact(oracle, modcod, F):
    from dvbs2_b200 import Decoder_BCH_DVBS2_cuda, Handle

    mc = MODCODS[modcod]
    K, N = mc["K_bch"], mc["N_bch"]
    U, Y = _bch_stress(oracle, K, N, F, seed=F)
    h = Handle(modcod=modcod, n_frames=F)
    dec = Decoder_BCH_DVBS2_cuda(h, F)
    V = np.zeros((F, K), np.int32)
    cwd = np.zeros(F, np.int8)
    dec.decode_hiho(Y, cwd, V)
    Vo, co, nc = oracle.bch_decode(K, N, Y)
    assert np.array_equal(cwd, co)
    assert np.array_equal(V, Vo)
    st = h.stats()
    nonzero = int((oracle.bch_decode(K, N, Y)[2] > 0).sum() + ((co == 0).sum()))
    assert st["bch_failures"] == int((co == 0).sum())
    assert st["bch_nonzero"] >= st["bch_failures"] and st["bch_nonzero"] <= F
    h.close()


def test_bch_random_words_are_rejected_like_the_reference(oracle):
    """Random words (hundreds of errors): Berlekamp must abort / Chien must disagree exactly as the oracle does."""
    from dvbs2_b200 import Decoder_BCH_DVBS2_cuda, Handle

    K, N, F = 9552, 9720, 64
    rng = np.random.default_rng(11)
    Y = rng.integers(0, 2, size=(F, N), dtype=np.int32)
    h = Handle(modcod="QPSK-S_3/5", n_frames=F)
    V = np.zeros((F, K), np.int32)
    cwd = np.zeros(F, np.int8)
    Decoder_BCH_DVBS2_cuda(h, F).decode_hiho(Y, cwd, V)
    Vo, co, _ = oracle.bch_decode(K, N, Y)
    assert np.array_equal(cwd, co) and np.array_equal(V, Vo)
    h.close()


@pytest.mark.parametrize("modcod,ebn0", list(zip(ALL, [4.0, 2.3, 4.0, 6.8, 7.8])))
def test_demodulate_and_deinterleave(oracle, modcod, ebn0):
    from dvbs2_b200 import Handle, Interleaver_DVBS2_cuda, Modem_DVBS2_cuda

    mc = MODCODS[modcod]
    F = 40
    s = synth_frames(oracle, modcod, F, ebn0)
    s["sigma"] = (s["sigma"] * np.linspace(0.8, 1.3, F)).astype(np.float32)  # per-frame sigma, as the estimator gives
    h = Handle(modcod=modcod, n_frames=F)
    llr = np.zeros((F, N_LDPC), np.float32)
    Modem_DVBS2_cuda(h, F).demodulate(s["sigma"], s["rx"], llr)
    want = oracle.demodulate(mc["bps"], s["sigma"], s["rx"])
    err = np.abs(llr - want) / np.maximum(1.0, np.abs(want))
    assert err.max() <= 1e-4, err.max()  # SURVEY App. A.1 parity criterion
    nat = np.zeros_like(llr)
    Interleaver_DVBS2_cuda(h, F).deinterleave(llr, nat)
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    assert np.array_equal(nat, llr[:, inv])
    h.close()


@pytest.mark.parametrize("prec,dec,bits", [("i16", 2, 6), ("i8", 2, 6), ("i8", 3, 8)])
def test_quantizer_bit_exact(oracle, prec, dec, bits):
    import ctypes as C

    from dvbs2_b200 import Handle

    rng = np.random.default_rng(5)
    y = (rng.standard_normal(100003) * 6).astype(np.float32)
    y[:8] = [0.125, -0.125, 0.375, -0.375, 1e9, -1e9, 0.0, 7.875]
    h = Handle(precision=prec, quant_bits=bits, quant_dec=dec, n_frames=32)
    out = np.zeros(y.shape, np.int16 if prec == "i16" else np.int8)
    h.check(h.lib.dvbs2_cuda_quantize(h._h, y.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), y.size))
    assert np.array_equal(out, oracle.quantize(y, dec, bits, prec))
    h.close()


def _oracle_chain(oracle, modcod, s, llr, prec="f32", dec=2, bits=6):
    mc = MODCODS[modcod]
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    nat = np.ascontiguousarray(llr[:, inv])
    if prec != "f32":
        nat = oracle.quantize(nat, dec, bits, prec)
    Vl, cl, nl = oracle.ldpc_decode(mc["N_bch"], nat, prec=prec)
    Vb, cb, _ = oracle.bch_decode(mc["K_bch"], mc["N_bch"], Vl)
    return Vl, cl, nl, Vb, cb


@pytest.mark.parametrize("nf", [0, 2])  # frames per lane (f32): chosen by the library (1 at this size) / the large-batch shape
@pytest.mark.parametrize("modcod,ebn0,prec", [("QPSK-S_8/9", 4.0, "f32"), ("QPSK-S_3/5", 2.4, "f32"),
                                              ("8PSK-S_3/5", 4.0, "i16"), ("8PSK-S_3/5", 4.2, "i8"),
                                              ("8PSK-S_8/9", 7.5, "f32"), ("16APSK-S_8/9", 7.8, "f32")])
def test_fused_chain_matches_oracle_chain(oracle, modcod, ebn0, prec, nf):
    """The chain on the device must equal oracle-LDPC + oracle-BCH run on the *GPU's own* LLRs (decoders are judged on
    identical LLRs; the demapper has its own tolerance test above)."""
    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle, Modem_DVBS2_cuda

    if prec != "f32" and nf:
        pytest.skip("the integer precisions have one group shape")
    mc = MODCODS[modcod]
    F = 96
    s = synth_frames(oracle, modcod, F, ebn0, seed=17)
    h = Handle(modcod=modcod, n_frames=F, precision=prec, ldpc_frames_per_lane=nf)
    llr = np.zeros((F, N_LDPC), np.float32)
    Modem_DVBS2_cuda(h, F).demodulate(s["sigma"], s["rx"], llr)
    V = np.zeros((F, mc["K_bch"]), np.int32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    Decoder_chain_DVBS2_cuda(h, F).demodulate_decode(s["sigma"], s["rx"], c1, c2, V, nit)
    Vl, cl, nl, Vb, cb = _oracle_chain(oracle, modcod, s, llr, prec)
    assert np.array_equal(nit, nl) and np.array_equal(c1, cl)
    assert np.array_equal(c2, cb) and np.array_equal(V, Vb)
    # and the payload comes back after BB descrambling wherever BCH says so
    good = c2 == 1
    assert np.array_equal(oracle.bb_scramble(mc["K_bch"], V[good]), s["U"][good])
    h.close()


def test_chain_divergent_iteration_counts(oracle):
    """BASELINE config 4: 8PSK-S_8/9 at high Eb/N0 — most frames stop after 1-3 iterations, a few run long."""
    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle, Modem_DVBS2_cuda

    modcod, F = "8PSK-S_8/9", 128
    mc = MODCODS[modcod]
    s = synth_frames(oracle, modcod, F, 7.5, seed=23)
    # make a few frames much noisier so iteration counts really diverge inside one warp
    rng = np.random.default_rng(1)
    for f in (3, 40, 77):
        s["rx"][f] += rng.standard_normal(s["rx"][f].shape).astype(np.float32) * 0.05
    h = Handle(modcod=modcod, n_frames=F, ldpc_frames_per_lane=2)
    llr = np.zeros((F, N_LDPC), np.float32)
    Modem_DVBS2_cuda(h, F).demodulate(s["sigma"], s["rx"], llr)
    V = np.zeros((F, mc["K_bch"]), np.int32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    Decoder_chain_DVBS2_cuda(h, F).demodulate_decode(s["sigma"], s["rx"], c1, c2, V, nit)
    Vl, cl, nl, Vb, cb = _oracle_chain(oracle, modcod, s, llr)
    assert np.array_equal(nit, nl) and np.array_equal(c1, cl) and np.array_equal(c2, cb) and np.array_equal(V, Vb)
    assert nit.min() < nit.max()
    h.close()


def test_device_resident_api_matches_host_api(oracle):
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle

    modcod, F = "QPSK-S_8/9", 160
    mc = MODCODS[modcod]
    s = synth_frames(oracle, modcod, F, 4.0, seed=31)
    h = Handle(modcod=modcod, n_frames=F)
    ch = Decoder_chain_DVBS2_cuda(h, F)
    V = np.zeros((F, mc["K_bch"]), np.int32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    ch.demodulate_decode(s["sigma"], s["rx"], c1, c2, V, nit)
    dev = torch.device("cuda:0")
    d_sig, d_rx = torch.from_numpy(s["sigma"]).to(dev), torch.from_numpy(s["rx"]).to(dev)
    dV = torch.zeros((F, mc["K_bch"]), dtype=torch.int32, device=dev)
    d1, d2 = torch.zeros(F, dtype=torch.int8, device=dev), torch.zeros(F, dtype=torch.int8, device=dev)
    dn = torch.zeros(F, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    ch.demodulate_decode_dev(d_sig, d_rx, d1, d2, dV, dn)
    h.sync()
    assert np.array_equal(dV.cpu().numpy(), V) and np.array_equal(d1.cpu().numpy(), c1)
    assert np.array_equal(d2.cpu().numpy(), c2) and np.array_equal(dn.cpu().numpy(), nit)
    h.close()


@pytest.mark.parametrize("modcod,ebn0,F", [("QPSK-S_8/9", 4.0, 70), ("8PSK-S_3/5", 4.0, 33)])
def test_packed_output_format(oracle, modcod, ebn0, F):
    """the bit-packed socket format is the int32 output packed like spu::tools::Bit_packer (LSB first), and
    with descramble it is the packed Scrambler_BB::descramble output; host and device entry points agree."""
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle

    mc = MODCODS[modcod]
    K = mc["K_bch"]
    s = synth_frames(oracle, modcod, F, ebn0, seed=53)
    h = Handle(modcod=modcod, n_frames=F)
    ch = Decoder_chain_DVBS2_cuda(h, F)
    V = np.zeros((F, K), np.int32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    ch.demodulate_decode(s["sigma"], s["rx"], c1, c2, V, nit)
    for descramble in (False, True):
        P = np.zeros((F, K // 8), np.uint8)
        p1, p2, pn = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
        ch.demodulate_decode_packed(s["sigma"], s["rx"], p1, p2, P, pn, descramble=descramble)
        want = oracle.bb_scramble(K, V) if descramble else V
        assert np.array_equal(P, np.packbits(want.astype(np.uint8), axis=1, bitorder="little"))
        assert np.array_equal(p1, c1) and np.array_equal(p2, c2) and np.array_equal(pn, nit)
        dev = torch.device("cuda:0")
        d_sig, d_rx = torch.from_numpy(s["sigma"]).to(dev), torch.from_numpy(s["rx"]).to(dev)
        dP = torch.zeros((F, K // 8), dtype=torch.uint8, device=dev)
        torch.cuda.synchronize()
        ch.demodulate_decode_packed_dev(d_sig, d_rx, None, None, dP, None, descramble=descramble)
        h.sync()
        assert np.array_equal(dP.cpu().numpy(), P)
    h.close()


def test_chain_full_size_round_trip(oracle):
    """BASELINE-size property test: 40960 frames through the chunked host API (two chunks on two pipeline streams: a call is
    split once it exceeds four groups per SM); every frame BCH accepts must descramble to its payload, and at 5 dB (well past
    the waterfall) every frame is accepted."""
    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle

    modcod, F0, reps = "QPSK-S_8/9", 64, 640
    mc = MODCODS[modcod]
    s = synth_frames(oracle, modcod, F0, 5.0, seed=41)
    F = F0 * reps
    rx = np.ascontiguousarray(np.tile(s["rx"], (reps, 1)))
    sig = np.ascontiguousarray(np.tile(s["sigma"], reps))
    h = Handle(modcod=modcod, n_frames=F)
    V = np.zeros((F, mc["K_bch"]), np.int32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    Decoder_chain_DVBS2_cuda(h, F).demodulate_decode(sig, rx, c1, c2, V, nit)
    assert (c2 == 1).all() and (c1 == 1).all()
    assert np.array_equal(oracle.bb_scramble(mc["K_bch"], V), np.tile(s["U"], (reps, 1)))
    assert np.array_equal(nit, np.tile(nit[:F0], reps))  # identical inputs -> identical iteration counts in any lane
    h.close()


def test_errors_are_reported_not_swallowed():
    from dvbs2_b200 import Decoder_LDPC_DVBS2_cuda, Handle, capi

    with pytest.raises(capi.Dvbs2CudaError):
        Handle(norm=0.3, precision="i16")  # integers only know k/8 factors (Update_rule_NMS.hxx:14-39)
    with pytest.raises(ValueError):
        Handle(modcod="32APSK-S_9/10")
    h = Handle(n_frames=32)
    dec = Decoder_LDPC_DVBS2_cuda(h, 32)
    with pytest.raises(ValueError):
        dec.decode_siho(np.zeros(10, np.float32), np.zeros(32, np.int8), np.zeros((32, dec.K), np.int32))
    with pytest.raises(ValueError):
        dec.set_n_frames(h.info["max_frames"] + 1)
    h.close()


def test_empty_batches_and_bad_arguments_through_the_c_abi():
    """Zero frames is a no-op that touches no buffer; too many frames, NULL buffers and ambiguous outputs are refused with the
    documented status codes (no exception crosses the ABI, nothing is silently clamped)."""
    import ctypes as C

    from dvbs2_b200 import Handle, capi

    h = Handle(modcod="QPSK-S_8/9", n_frames=64)
    L, K = h.lib, h.info["K_bch"]
    cp = np.full(64, 0.3, np.float32)
    sym = np.zeros((64, 16200), np.float32)
    V = np.full((64, K), 7, np.int32)
    cw = np.full(64, 7, np.int8)
    P = np.full((64, K // 8), 7, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    launches = h.stats()["kernel_launches"]
    assert L.dvbs2_cuda_chain(h._h, p(cp), p(sym), p(cw), p(cw), p(V), None, 0) == capi.OK
    assert L.dvbs2_cuda_chain_packed(h._h, p(cp), p(sym), None, None, p(P), None, 1, 0) == capi.OK
    assert L.dvbs2_cuda_ldpc_decode_siho(h._h, p(sym), p(cw), p(V), None, 0) == capi.OK
    assert L.dvbs2_cuda_bch_decode_hiho(h._h, p(V), p(cw), p(V), 0) == capi.OK
    assert L.dvbs2_cuda_demodulate(h._h, p(cp), p(sym), p(sym), 0) == capi.OK
    assert L.dvbs2_cuda_rx_chain(h._h, p(sym), None, None, None, None, p(V), None, 1, None, 0) == capi.OK
    assert (V == 7).all() and (cw == 7).all() and (P == 7).all() and h.stats()["kernel_launches"] == launches
    too_many = h.info["max_frames"] + 1
    assert L.dvbs2_cuda_chain(h._h, p(cp), p(sym), None, None, p(V), None, too_many) == capi.ERR_TOO_MANY
    assert L.dvbs2_cuda_chain(h._h, p(cp), p(sym), None, None, p(V), None, -1) == capi.ERR_INVALID
    assert L.dvbs2_cuda_chain(h._h, None, p(sym), None, None, p(V), None, 64) == capi.ERR_INVALID
    assert L.dvbs2_cuda_chain(h._h, p(cp), p(sym), None, None, None, None, 64) == capi.ERR_INVALID
    assert L.dvbs2_cuda_chain_packed(h._h, p(cp), p(sym), None, None, None, None, 0, 64) == capi.ERR_INVALID
    # rx_chain wants exactly one output format
    pl = np.zeros((64, h.lib.dvbs2_cuda_pl_frame_size(h._h)), np.float32)
    assert L.dvbs2_cuda_rx_chain(h._h, p(pl), None, None, None, None, p(V), p(P), 1, None, 64) == capi.ERR_INVALID
    assert L.dvbs2_cuda_rx_chain(h._h, p(pl), None, None, None, None, None, None, 1, None, 64) == capi.ERR_INVALID
    assert (V == 7).all()
    h.close()
