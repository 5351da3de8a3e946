"""GPU parity: LDPC decoder through the C ABI vs the CPU oracle on identical LLRs (bit-exact, all precisions)."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, synth_frames

pytestmark = pytest.mark.gpu


_CACHE = {}


def _llrs(oracle, modcod, F, ebn0, seed=1):
    s = synth_frames(oracle, modcod, F, ebn0, seed=seed)
    mc = MODCODS[modcod]
    llr = oracle.demodulate(mc["bps"], s["sigma"], s["rx"])
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    return np.ascontiguousarray(llr[:, inv]), s


def _decode(modcod, Y, prec="f32", n_ite=10, norm=1.0, early_stop=True, compaction=True, **cfg):
    """Float calls run three times unless the caller pins it: through the cooperative small-batch kernel (what the library
    picks at these sizes), through the lane = frame kernel with 32-frame groups, and with 64-frame groups (the large-batch
    shape); all must give the same bits, CWD and iteration counts."""
    if prec == "f32" and "ldpc_frames_per_lane" not in cfg and "ldpc_onchip" not in cfg:
        a = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_frames_per_lane=1, **cfg)
        b = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_frames_per_lane=1, ldpc_coop_max_frames=-1, **cfg)
        c = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_frames_per_lane=2, **cfg)
        d = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_onchip=1, **cfg)  # design B: variables on chip
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], c[:3])) and all(np.array_equal(x, y) for x, y in zip(b[:3], c[:3]))
        assert all(np.array_equal(x, y) for x, y in zip(d[:3], c[:3]))
        for k in ("frames", "ldpc_iterations", "ldpc_failures", "iter_hist"):
            assert d[3][k] == c[3][k], k
        return c
    if prec == "i16" and "ldpc_coop_max_frames" not in cfg:
        # int16 calls of this size go to the cooperative small-batch kernel (one frame per lane); the packed-halfword lane = frame
        # kernel (what large batches use) must give the same bits, CWD, iteration counts and statistics
        a = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_coop_max_frames=0, **cfg)
        b = _decode(modcod, Y, prec, n_ite, norm, early_stop, compaction, ldpc_coop_max_frames=-1, **cfg)
        assert all(np.array_equal(x, y) for x, y in zip(a[:3], b[:3]))
        for k in ("frames", "ldpc_iterations", "ldpc_failures", "iter_hist"):
            assert a[3][k] == b[3][k], k
        return b
    from dvbs2_b200 import Decoder_LDPC_DVBS2_cuda, Handle

    F = Y.shape[0]
    h = Handle(modcod=modcod, n_ite=n_ite, norm=norm, precision=prec, early_stop=early_stop, n_frames=F,
               compaction=compaction, **cfg)
    dec = Decoder_LDPC_DVBS2_cuda(h, F)
    V = np.zeros((F, dec.K), np.int32)
    cwd = np.zeros(F, np.int8)
    nit = np.zeros(F, np.int32)
    dec.decode_siho(Y, cwd, V, nit)
    st = h.stats()
    h.close()
    return V, cwd, nit, st


@pytest.mark.parametrize("modcod,ebn0", [("QPSK-S_8/9", 3.9), ("QPSK-S_3/5", 2.3), ("8PSK-S_3/5", 4.0),
                                         ("8PSK-S_8/9", 6.7), ("16APSK-S_8/9", 7.7)])
def test_f32_bit_exact_all_modcods(oracle, modcod, ebn0):
    Y, _ = _llrs(oracle, modcod, 96, ebn0)
    K = MODCODS[modcod]["N_bch"]
    V, cwd, nit, st = _decode(modcod, Y)
    Vo, co, no = oracle.ldpc_decode(K, Y)
    assert np.array_equal(nit, no)
    assert np.array_equal(cwd, co)
    assert np.array_equal(V, Vo)
    assert st["frames"] == 96 and st["ldpc_iterations"] == int(no.sum()) and st["ldpc_failures"] == int((co == 0).sum())


@pytest.mark.parametrize("F", [1, 31, 32, 33, 70])
def test_ragged_batches(oracle, F):
    Y, _ = _llrs(oracle, "QPSK-S_8/9", F, 4.0, seed=F)
    V, cwd, nit, _ = _decode("QPSK-S_8/9", Y)
    Vo, co, no = oracle.ldpc_decode(14400, Y)
    assert np.array_equal(V, Vo) and np.array_equal(cwd, co) and np.array_equal(nit, no)


@pytest.mark.parametrize("prec,dec,bits", [("i16", 2, 6), ("i8", 2, 6), ("i8", 1, 6), ("i16", 3, 8)])
@pytest.mark.parametrize("modcod,ebn0", [("QPSK-S_8/9", 4.0), ("8PSK-S_3/5", 4.0)])
def test_fixed_point_bit_exact(oracle, prec, dec, bits, modcod, ebn0):
    Y, _ = _llrs(oracle, modcod, 64, ebn0)
    q = oracle.quantize(Y, dec, bits, prec)
    K = MODCODS[modcod]["N_bch"]
    V, cwd, nit, _ = _decode(modcod, q, prec=prec)
    Vo, co, no = oracle.ldpc_decode(K, q, prec=prec)
    assert np.array_equal(nit, no) and np.array_equal(cwd, co) and np.array_equal(V, Vo)


@pytest.mark.parametrize("norm", [0.75, 0.875, 0.5])
@pytest.mark.parametrize("prec", ["f32", "i16"])
def test_normalisation_factors(oracle, norm, prec):
    Y, _ = _llrs(oracle, "QPSK-S_8/9", 64, 3.9)
    if prec != "f32":
        Y = oracle.quantize(Y, 2, 6, prec)
    V, cwd, nit, _ = _decode("QPSK-S_8/9", Y, prec=prec, norm=norm)
    Vo, co, no = oracle.ldpc_decode(14400, Y, norm=norm, prec=prec)
    assert np.array_equal(nit, no) and np.array_equal(cwd, co) and np.array_equal(V, Vo)


@pytest.mark.parametrize("modcod", ["QPSK-S_8/9", "QPSK-S_3/5"])
def test_int16_wraps_like_the_reference(oracle, modcod):
    """int16 inputs far outside the quantiser's range make the reference's int16_t arithmetic overflow and wrap; the packed
    halfword kernel must wrap identically (ring arithmetic modulo 2^16 per frame), norm 1 and k/8 factors alike."""
    K = MODCODS[modcod]["N_bch"]
    Y, _ = _llrs(oracle, modcod, 130, 4.0 if K == 14400 else 2.4, seed=9)
    rng = np.random.default_rng(12)
    q = np.clip(np.round(Y * 900.0), -32767, 32767).astype(np.int16)            # frames 0..: large, often saturated
    q[40:80] = rng.integers(-32768, 32768, size=(40, N_LDPC)).astype(np.int16)    # pure noise over the full range
    q[80:100] = np.clip(np.round(Y[80:100] * 4.0), -31, 31).astype(np.int16)      # the default 6-bit quantiser
    q[100] = 0
    q[101] = -32768
    q[102] = 32767
    for norm in (1.0, 0.75):
        V, cwd, nit, _ = _decode(modcod, q, prec="i16", norm=norm)
        Vo, co, no = oracle.ldpc_decode(K, q, prec="i16", norm=norm)
        assert np.array_equal(nit, no) and np.array_equal(cwd, co) and np.array_equal(V, Vo)


def test_no_early_stop_and_iteration_counts(oracle):
    Y, _ = _llrs(oracle, "QPSK-S_8/9", 64, 4.2)
    V, cwd, nit, _ = _decode("QPSK-S_8/9", Y, n_ite=5, early_stop=False)
    Vo, co, no = oracle.ldpc_decode(14400, Y, n_ite=5, early_stop=False)
    assert (nit == 5).all() and np.array_equal(V, Vo) and np.array_equal(cwd, co)
    for n_ite in (1, 3, 20):
        V, cwd, nit, _ = _decode("QPSK-S_8/9", Y, n_ite=n_ite)
        Vo, co, no = oracle.ldpc_decode(14400, Y, n_ite=n_ite)
        assert np.array_equal(nit, no) and np.array_equal(V, Vo) and np.array_equal(cwd, co)


def test_extreme_inputs(oracle):
    """all-zero LLRs, huge LLRs, noiseless codewords, all-erased parity: no NaN surprises, still bit-exact."""
    rng = np.random.default_rng(3)
    s = synth_frames(oracle, "QPSK-S_8/9", 8, 100.0)
    clean = (1.0 - 2.0 * s["ldpc"].astype(np.float32)) * 8.0
    Y = np.stack([np.zeros(N_LDPC, np.float32), clean[1] * 1e30, clean[2], clean[3] * 1e-30,
                  np.where(np.arange(N_LDPC) >= 14400, 0.0, clean[4]).astype(np.float32),
                  rng.standard_normal(N_LDPC).astype(np.float32), -clean[6], clean[7]])
    Y = np.ascontiguousarray(Y, np.float32)
    V, cwd, nit, _ = _decode("QPSK-S_8/9", Y)
    Vo, co, no = oracle.ldpc_decode(14400, Y)
    assert np.array_equal(nit, no) and np.array_equal(cwd, co) and np.array_equal(V, Vo)


def test_large_batch_round_trip(oracle):
    """Size-independent property at a size the oracle would not finish quickly: noiseless-ish codewords decode to the
    transmitted bits in one iteration, for every frame of a multi-thousand batch."""
    s = synth_frames(oracle, "QPSK-S_8/9", 16, 12.0)
    mc = MODCODS["QPSK-S_8/9"]
    llr = oracle.demodulate(2, s["sigma"], s["rx"])
    reps = 160  # 2560 frames
    Y = np.ascontiguousarray(np.tile(llr, (reps, 1)))
    V, cwd, nit, st = _decode("QPSK-S_8/9", Y)
    assert (cwd == 1).all() and (nit == 1).all()
    assert np.array_equal(V, np.tile(s["bch"], (reps, 1)))
    assert st["frames"] == 16 * reps


@pytest.mark.parametrize("modcod,ebn0,prec,F", [("QPSK-S_8/9", 4.0, "f32", 330), ("QPSK-S_8/9", 4.1, "i16", 200),
                                                ("8PSK-S_3/5", 4.1, "f32", 200), ("QPSK-S_8/9", 4.0, "i8", 300),
                                                ("8PSK-S_8/9", 7.0, "f32", 257)])
def test_frame_compaction_is_transparent(oracle, modcod, ebn0, prec, F):
    """Stragglers are moved to dense groups for their remaining iterations: bits, CWD, iteration counts and counters must
    equal both the in-place decoder (compaction off) and the oracle."""
    Y, _ = _llrs(oracle, modcod, F, ebn0, seed=77)
    if prec != "f32":
        Y = oracle.quantize(Y, 1 if prec == "i8" else 2, 6, prec)
    kw = dict(prec=prec)
    K = MODCODS[modcod]["N_bch"]
    V1, c1, n1, s1 = _decode(modcod, Y, compaction=True, **kw) if prec != "i8" else _decode_i8(modcod, Y, True)
    V0, c0, n0, s0 = _decode(modcod, Y, compaction=False, **kw) if prec != "i8" else _decode_i8(modcod, Y, False)
    assert np.array_equal(V1, V0) and np.array_equal(c1, c0) and np.array_equal(n1, n0)
    for k in ("frames", "ldpc_iterations", "ldpc_failures"):
        assert s1[k] == s0[k], k
    Vo, co, no = oracle.ldpc_decode(K, Y, prec=prec)
    assert np.array_equal(V1, Vo) and np.array_equal(c1, co) and np.array_equal(n1, no)
    assert n1.min() < n1.max()  # iteration counts really diverge, so frames were handed over
    assert s1["kernel_launches"] > s0["kernel_launches"]


@pytest.mark.parametrize("rounds,eighths,nf_later", [(3, 4, 1), (4, 4, 2), (3, 5, 2), (4, 6, 1), (2, 3, 2), (3, 2, 1), (6, 6, 2),
                                                     (5, 7, 1), (6, 7, 2)])
def test_multi_round_compaction_is_transparent(oracle, rounds, eighths, nf_later):
    """Any number of rounds, hand-over threshold and later-round group width gives the oracle's bits, CWD and counts
    (frames handed over twice or three times find their way back to their own bit lane)."""
    modcod, F = "QPSK-S_8/9", 1100
    if "mr" not in _CACHE:
        Y, _ = _llrs(oracle, modcod, F, 4.0, seed=91)
        _CACHE["mr"] = (Y,) + tuple(oracle.ldpc_decode(MODCODS[modcod]["N_bch"], Y))
    Y, Vo, co, no = _CACHE["mr"]
    V, c, n, st = _decode(modcod, Y, compaction_rounds=rounds, compaction_exit_eighths=eighths, compaction_nf_later=nf_later)
    assert np.array_equal(n, no) and np.array_equal(c, co) and np.array_equal(V, Vo)
    assert st["frames"] == F and st["ldpc_iterations"] == int(no.sum()) and st["ldpc_failures"] == int((co == 0).sum())
    assert st["kernel_launches"] >= 1 + 4 * (rounds - 1)


@pytest.mark.parametrize("prec,modcod,ebn0", [("i16", "8PSK-S_3/5", 4.0), ("i8", "QPSK-S_8/9", 4.0)])
def test_multi_round_compaction_fixed_point(oracle, prec, modcod, ebn0):
    F = 700
    Y, _ = _llrs(oracle, modcod, F, ebn0, seed=92)
    Y = oracle.quantize(Y, 1 if prec == "i8" else 2, 6, prec)
    from dvbs2_b200 import Decoder_LDPC_DVBS2_cuda, Handle

    h = Handle(modcod=modcod, precision=prec, quant_dec=1 if prec == "i8" else 2, n_frames=F, compaction_rounds=3)
    dec = Decoder_LDPC_DVBS2_cuda(h, F)
    V, cwd, nit = np.zeros((F, dec.K), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int32)
    dec.decode_siho(Y, cwd, V, nit)
    h.close()
    Vo, co, no = oracle.ldpc_decode(MODCODS[modcod]["N_bch"], Y, prec=prec)
    assert np.array_equal(nit, no) and np.array_equal(cwd, co) and np.array_equal(V, Vo)


def _decode_i8(modcod, Y, compaction):
    from dvbs2_b200 import Decoder_LDPC_DVBS2_cuda, Handle

    F = Y.shape[0]
    h = Handle(modcod=modcod, precision="i8", quant_dec=1, n_frames=F, compaction=compaction)
    dec = Decoder_LDPC_DVBS2_cuda(h, F)
    V, cwd, nit = np.zeros((F, dec.K), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int32)
    dec.decode_siho(Y, cwd, V, nit)
    st = h.stats()
    h.close()
    return V, cwd, nit, st
