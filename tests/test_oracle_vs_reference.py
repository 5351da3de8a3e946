"""CPU suite: the oracle against the compiled, unmodified reference (oracle/_ref/libdvbs2_ref.so) on seeded inputs.
Skipped where the reference was never built; the golden fixtures (test_oracle_golden.py) cover those hosts."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, synth_frames

ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]
EBN0 = dict(zip(ALL, [3.95, 2.35, 3.9, 6.7, 7.7]))


def _nat_llrs(oracle, ref, modcod, F, seed):
    s = synth_frames(oracle, modcod, F, EBN0[modcod], seed=seed)
    _, nat, _ = ref.demodulate(modcod, s["sigma"], s["rx"])
    return s, nat


def test_H_and_field_tables(oracle, ref):
    for K in (14400, 9720):
        off, cols = oracle.H_arrays(K)
        deg, rc = ref.build_H(K)
        assert np.array_equal(np.diff(off), deg) and np.array_equal(cols, rc)
    for a, b in zip(oracle.gf_tables(), ref.bch_tables()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("modcod", ALL)
def test_tx_and_demapper(oracle, ref, modcod):
    mc = MODCODS[modcod]
    rng = np.random.default_rng(3)
    U = rng.integers(0, 2, size=(8, mc["K_bch"]), dtype=np.int32)
    a, b = oracle.tx(modcod, U), ref.tx(modcod, U)
    for k in ("scr", "bch", "ldpc", "itl", "sym"):
        assert np.array_equal(a[k], b[k]), k
    s = synth_frames(oracle, modcod, 8, EBN0[modcod], seed=4)
    itl, nat, cst = ref.demodulate(modcod, s["sigma"], s["rx"])
    mine = oracle.demodulate(mc["bps"], s["sigma"], s["rx"])
    assert (np.abs(mine - itl) / np.maximum(1, np.abs(itl))).max() <= 1e-4
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    assert np.array_equal(itl[:, inv], nat) and np.array_equal(oracle.constellation(mc["bps"]), cst)


@pytest.mark.parametrize("modcod", ALL)
def test_ldpc_float_sequential_and_inter(oracle, ref, modcod):
    N = MODCODS[modcod]["N_bch"]
    _, nat = _nat_llrs(oracle, ref, modcod, 64, seed=11)
    V, cwd, _ = oracle.ldpc_decode(N, nat)
    Vs, cs = ref.ldpc_decode(N, nat, simd=False)
    assert np.array_equal(cwd, cs) and np.array_equal(V, Vs)
    Vi, ci = ref.ldpc_decode(N, nat, simd=True)
    ok = ci == 1
    assert ok.sum() > 8 and np.array_equal(V[ok], Vi[ok]) and (cwd[ok] == 1).all()
    # frames/CWD may differ from the wave-level-stop SIMD decoder only on frames that did not converge there
    assert abs(int((cwd == 0).sum()) - int((ci == 0).sum())) <= 2


@pytest.mark.parametrize("modcod", ["QPSK-S_8/9", "8PSK-S_3/5"])
@pytest.mark.parametrize("prec,dec,bits", [("i16", 2, 6), ("i16", 3, 8), ("i8", 2, 6), ("i8", 1, 6)])
def test_ldpc_fixed_point(oracle, ref, modcod, prec, dec, bits):
    N = MODCODS[modcod]["N_bch"]
    _, nat = _nat_llrs(oracle, ref, modcod, 32, seed=12)
    q = ref.quantize(nat, dec, bits, prec)
    assert np.array_equal(q, oracle.quantize(nat, dec, bits, prec))
    V, cwd, _ = oracle.ldpc_decode(N, q, prec=prec)
    Vs, cs = ref.ldpc_decode(N, q, prec=prec, simd=False)
    assert np.array_equal(cwd, cs) and np.array_equal(V, Vs)
    if prec == "i16":
        Vi, ci = ref.ldpc_decode(N, q, prec=prec, simd=True)
        ok = ci == 1
        assert np.array_equal(V[ok], Vi[ok])


@pytest.mark.parametrize("norm", [0.75, 0.5, 0.875])
def test_ldpc_normalisation_factors(oracle, ref, norm):
    _, nat = _nat_llrs(oracle, ref, "QPSK-S_8/9", 32, seed=13)
    V, cwd, _ = oracle.ldpc_decode(14400, nat, norm=norm)
    Vs, cs = ref.ldpc_decode(14400, nat, norm=norm, simd=False)
    assert np.array_equal(cwd, cs) and np.array_equal(V, Vs)
    q = ref.quantize(nat, 2, 6, "i16")
    V, cwd, _ = oracle.ldpc_decode(14400, q, norm=norm, prec="i16")
    Vs, cs = ref.ldpc_decode(14400, q, norm=norm, prec="i16", simd=False)
    assert np.array_equal(cwd, cs) and np.array_equal(V, Vs)


@pytest.mark.parametrize("K,N", [(14232, 14400), (9552, 9720)])
def test_bch(oracle, ref, K, N):
    rng = np.random.default_rng(N)
    U = rng.integers(0, 2, size=(48, K), dtype=np.int32)
    Y = oracle.bch_encode(K, N, U)
    for f in range(48):
        ne = [0, 1, 3, 11, 12, 13, 14, 12, 25, 200, 12, 13][f % 12]
        Y[f, rng.choice(N, size=ne, replace=False)] ^= 1
    V, cwd, _ = oracle.bch_decode(K, N, Y)
    Vr, cr = ref.bch_decode(K, N, Y)
    assert np.array_equal(cwd, cr) and np.array_equal(V, Vr)
    assert np.array_equal(oracle.bb_scramble(K, V), ref.descramble(K, V))


@pytest.mark.parametrize("modcod", ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"])
def test_framer_and_estimator(oracle, ref, modcod):
    """Framer::remove_plh (exact) and Estimator_DVBS2::estimate (FMA contraction in the compiled reference: 3e-6)."""
    from oracle.oracle import MODCODS, N_LDPC, synth_frames

    mc = MODCODS[modcod]
    n_sym = N_LDPC // mc["bps"]
    pl_sym = oracle.pl_frame_symbols(n_sym)
    s = synth_frames(oracle, modcod, 5, 5.0, seed=3)
    pl = ref.framer_generate(modcod, n_sym, pl_sym, s["rx"])
    a, b = oracle.remove_plh(n_sym, pl), ref.remove_plh(modcod, n_sym, pl_sym, pl)
    assert np.array_equal(a, b) and np.array_equal(a, s["rx"])
    eo, er = oracle.estimate(n_sym, mc["K_bch"], mc["bps"], a), ref.estimate(n_sym, mc["K_bch"], mc["bps"], a)
    assert np.max(np.abs(eo[0] - er[0]) / er[0]) < 3e-6
    assert np.allclose(eo[1], er[1], rtol=0, atol=1e-4) and np.allclose(eo[2], er[2], rtol=0, atol=1e-4)


@pytest.mark.parametrize("modcod", ["QPSK-S_8/9", "QPSK-S_3/5"])
def test_ldpc_templated_inter_decoder(oracle, ref, modcod):
    """Decoder_LDPC_BP_horizontal_layered_inter<B,R,Update_rule_NMS_simd<R>> (what AFF3CT's factory builds
    for --dec-simd INTER in C++17 builds) against the oracle, float and int16: identical on every frame it reports as
    converged, and identical to ONMS_inter (a8) altogether."""
    N = MODCODS[modcod]["N_bch"]
    _, nat = _nat_llrs(oracle, ref, modcod, 64, seed=21)
    V, cwd, _ = oracle.ldpc_decode(N, nat)
    Vt, ct = ref.ldpc_decode(N, nat, simd="tinter")
    Vi, ci = ref.ldpc_decode(N, nat, simd=True)
    ok = ct == 1
    assert ok.sum() > 8 and np.array_equal(V[ok], Vt[ok]) and (cwd[ok] == 1).all()
    assert np.array_equal(Vt, Vi) and np.array_equal(ct, ci)
    for norm in (1.0, 0.75):
        q = ref.quantize(nat, 2, 6, "i16")
        Vq, cq, _ = oracle.ldpc_decode(N, q, prec="i16", norm=norm)
        Vt, ct = ref.ldpc_decode(N, q, prec="i16", simd="tinter", norm=norm)
        ok = ct == 1
        assert ok.sum() > 8 and np.array_equal(Vq[ok], Vt[ok])


def test_ldpc_without_syndrome_reports_every_frame_valid(oracle, ref):
    """enable_syndrome = false: the reference never tests the syndrome, returns status 0 and CWD = 1 for every frame
    (Decoder_LDPC_BP.hxx:8-20); the oracle (and the GPU path with early_stop = 0) mirror that."""
    _, nat = _nat_llrs(oracle, ref, "QPSK-S_8/9", 16, seed=22)
    Vr, cr = ref.ldpc_decode(14400, nat, n_ite=4, enable_syndrome=False)
    V, cwd, nit = oracle.ldpc_decode(14400, nat, n_ite=4, early_stop=False)
    assert (cr == 1).all() and np.array_equal(cwd, cr) and np.array_equal(V, Vr) and (nit == 4).all()


def test_multithreaded_reference_drivers_equal_the_single_threaded_ones(oracle, ref):
    _, nat = _nat_llrs(oracle, ref, "QPSK-S_8/9", 48, seed=23)
    a = ref.ldpc_decode(14400, nat)
    b = ref.ldpc_decode(14400, nat, threads=5)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    c = ref.bch_decode(14232, 14400, a[0][:6], threads=2)
    d = ref.bch_decode(14232, 14400, a[0][:6])
    assert np.array_equal(c[0], d[0]) and np.array_equal(c[1], d[1])
