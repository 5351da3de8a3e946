"""CPU suite: the oracle (oracle/dvbs2_oracle.c) against the golden vectors generated from the compiled, unmodified
reference (tests/golden/make_golden.py) and against the reference's own shipped BER/FER trace for this decoder."""
import hashlib
import os

import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def ub(packed, n):
    return np.unpackbits(packed, axis=-1)[..., :n].astype(np.int32)


def load(modcod):
    return np.load(os.path.join(GOLD, "chain_%s.npz" % modcod.replace("/", "_")))


def test_tables_match_reference(oracle):
    t = np.load(os.path.join(GOLD, "tables.npz"))
    for K in (14400, 9720):
        off, cols = oracle.H_arrays(K)
        deg = np.diff(off).astype(np.int32)
        assert [len(deg), len(cols)] == list(t["H%d_M_E" % K])
        assert sha(deg) == str(t["H%d_deg_sha" % K]) and sha(cols.astype(np.int32)) == str(t["H%d_cols_sha" % K])
        assert np.array_equal(cols[: int(deg[:4].sum())], t["H%d_first_rows" % K])
        assert np.array_equal(cols[-int(deg[-1]):], t["H%d_last_row" % K])
    a, i, g = oracle.gf_tables()
    assert np.array_equal(a, t["gf_alpha_to"]) and np.array_equal(i, t["gf_index_of"]) and np.array_equal(g, t["bch_g"])
    assert len(g) == 169 and g[0] == 1 and g[-1] == 1
    for bps in (2, 3, 4):
        assert np.array_equal(oracle.constellation(bps), t["cstl_%d" % bps])  # bit-identical normalised points
    for name in ALL:
        mc = MODCODS[name]
        if mc["itl"]:
            rng = np.random.default_rng(1)
            U = rng.integers(0, 2, size=(1, mc["K_bch"]), dtype=np.int32)
            assert sha(oracle.tx(name, U)["itl"]) == str(t["itl_%s_sha" % name])


@pytest.mark.parametrize("modcod", ALL)
def test_tx_chain_matches_reference(oracle, modcod):
    g, mc = load(modcod), MODCODS[modcod]
    K, N = mc["K_bch"], mc["N_bch"]
    U = ub(g["U"], K)
    tx = oracle.tx(modcod, U)
    assert np.array_equal(tx["scr"], ub(g["scr"], K))
    assert np.array_equal(tx["bch"], ub(g["bch"], N))
    assert np.array_equal(tx["ldpc"], ub(g["ldpc"], N_LDPC))
    assert np.array_equal(tx["itl"], ub(g["itl"], N_LDPC))
    assert np.array_equal(tx["sym"], g["sym"])


@pytest.mark.parametrize("modcod", ALL)
def test_demapper_and_deinterleaver_match_reference(oracle, modcod):
    g, mc = load(modcod), MODCODS[modcod]
    llr = oracle.demodulate(mc["bps"], g["sigma"], g["rx"])
    # the reference's AVX body evaluates exp/log with MIPP polynomials: float-rounding parity
    err = np.abs(llr - g["llr_itl"]) / np.maximum(1.0, np.abs(g["llr_itl"]))
    assert err.max() <= 1e-4, err.max()
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    assert np.array_equal(g["llr_itl"][:, inv], g["llr_nat"])


@pytest.mark.parametrize("modcod", ALL)
def test_ldpc_float_matches_reference(oracle, modcod):
    g, mc = load(modcod), MODCODS[modcod]
    N = mc["N_bch"]
    V, cwd, nit = oracle.ldpc_decode(N, g["llr_nat"])
    assert np.array_equal(cwd, g["ldpc_seq_f32_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_f32_V"], N))
    assert 0 < cwd.sum() < len(cwd)  # the fixture holds both converged and failed frames
    # inter-frame SIMD decoder (what --dec-simd INTER builds): identical wherever it converged (north star criterion)
    ok = g["ldpc_inter_f32_cwd"] == 1
    assert ok.any() and np.array_equal(V[ok], ub(g["ldpc_inter_f32_V"], N)[ok]) and (cwd[ok] == 1).all()


@pytest.mark.parametrize("modcod", ALL)
def test_quantizer_and_fixed_point_ldpc_match_reference(oracle, modcod):
    g, mc = load(modcod), MODCODS[modcod]
    N = mc["N_bch"]
    assert np.array_equal(oracle.quantize(g["llr_nat"], 2, 6, "i16"), g["q16"])
    assert np.array_equal(oracle.quantize(g["llr_nat"], 2, 6, "i8"), g["q8"])
    assert np.array_equal(oracle.quantize(g["llr_nat"], 1, 6, "i8"), g["q8s1"])
    V, cwd, _ = oracle.ldpc_decode(N, g["q16"], prec="i16")
    assert np.array_equal(cwd, g["ldpc_seq_i16_cwd"]) and np.array_equal(V, ub(g["ldpc_seq_i16_V"], N))
    ok = g["ldpc_inter_i16_cwd"] == 1  # saturating SIMD int16 == wrapping sequential int16 with 6-bit inputs
    assert np.array_equal(V[ok], ub(g["ldpc_inter_i16_V"], N)[ok])
    for q, tag in (("q8", "i8"), ("q8s1", "i8s1")):  # int8 = wrapping sequential arithmetic, also when it fails
        V, cwd, _ = oracle.ldpc_decode(N, g[q], prec="i8")
        assert np.array_equal(cwd, g["ldpc_seq_%s_cwd" % tag]) and np.array_equal(V, ub(g["ldpc_seq_%s_V" % tag], N))


@pytest.mark.parametrize("modcod", ALL)
def test_bch_and_descrambler_match_reference(oracle, modcod):
    g, mc = load(modcod), MODCODS[modcod]
    K, N = mc["K_bch"], mc["N_bch"]
    V, cwd, _ = oracle.bch_decode(K, N, ub(g["ldpc_seq_f32_V"], N))
    assert np.array_equal(cwd, g["bch_cwd"]) and np.array_equal(V, ub(g["bch_V"], K))
    assert np.array_equal(oracle.bb_scramble(K, V), ub(g["desc"], K))
    good = cwd == 1
    assert np.array_equal(oracle.bb_scramble(K, V)[good], ub(g["U"], K)[good])


def test_bch_stress_vectors(oracle):
    g = np.load(os.path.join(GOLD, "bch_stress.npz"))
    for K, N in ((14232, 14400), (9552, 9720)):
        Y = ub(g["Y_%d" % K], N)
        V, cwd, ncorr = oracle.bch_decode(K, N, Y)
        assert np.array_equal(cwd, g["cwd_%d" % K]) and np.array_equal(V, ub(g["V_%d" % K], K))
        ne = g["nerr_%d" % K]
        assert (cwd[ne <= 12] == 1).all() and (cwd[ne > 12] == 0).any()


def test_nms_075_bpsk_curve_of_the_reference(oracle):
    """Known-answer test against the reference's shipped trace for this very decoder:
    lib/aff3ct/refs/LDPC/AWGN/BPSK/BP_HORIZONTAL_LAYERED/NMS/LDPC_N16200_K14400_hlayered_NMS_i10_DVB-S2.txt
    (16200,14400) NMS norm 0.75, 10 it, BPSK: FER 0.665 @ 3.6 dB (351/528), 0.276 @ 3.7 dB (212/768), 0.0717 @ 3.8 dB
    (128/1784).  Same check as the reference's CI: agreement within the joint binomial confidence interval."""
    K, R = 14400, 14400 / 16200
    rng = np.random.default_rng(2018)
    for ebn0, fe_ref, n_ref, F in ((3.6, 351, 528, 128), (3.7, 212, 768, 224), (3.8, 128, 1784, 256)):
        U = rng.integers(0, 2, size=(F, K), dtype=np.int32)
        X = oracle.ldpc_encode(K, U)
        sigma = np.sqrt(1.0 / (2.0 * 10 ** ((ebn0 + 10 * np.log10(R)) / 10)))
        y = (1.0 - 2.0 * X).astype(np.float32) + rng.standard_normal(X.shape, dtype=np.float32) * np.float32(sigma)
        llr = (y * np.float32(2.0 / sigma**2)).astype(np.float32)
        V, cwd, _ = oracle.ldpc_decode(K, llr, n_ite=10, norm=0.75)
        p_ref, p = fe_ref / n_ref, float((cwd == 0).mean())
        sd = np.sqrt(p_ref * (1 - p_ref) * (1.0 / n_ref + 1.0 / F))
        assert abs(p - p_ref) <= 4 * sd, (ebn0, p, p_ref, sd)
        # decoded words of converged frames are the transmitted ones
        assert np.array_equal(V[cwd == 1], U[cwd == 1])


@pytest.mark.parametrize("modcod", ALL)
def test_framer_and_estimator_match_reference(oracle, modcod):
    """oracle_remove_plh == Framer::remove_plh (exact) and oracle_estimate == Estimator_DVBS2::estimate (the
    compiled reference contracts a*a + b*b into FMAs, hence 3e-6 relative) on PL frames made by the reference's Framer::generate
    (tests/golden/make_golden_rx_front.py)."""
    g = np.load(os.path.join(GOLD, "rx_front.npz"))
    k, mc = modcod.replace("/", "_"), MODCODS[modcod]
    n_sym = N_LDPC // mc["bps"]
    pl = g[k + "_pl"]
    assert pl.shape[1] == 2 * oracle.pl_frame_symbols(n_sym)
    x = oracle.remove_plh(n_sym, pl)
    assert hashlib.sha256(np.ascontiguousarray(x).tobytes()).digest() == g[k + "_xfec_sha"].tobytes()
    assert np.array_equal(oracle.remove_plh(n_sym, oracle.insert_plh(n_sym, x)), x)
    sig, eb, es = oracle.estimate(n_sym, mc["K_bch"], mc["bps"], x)
    ref = g[k + "_est"]
    assert np.max(np.abs(sig - ref[0]) / ref[0]) < 3e-6
    assert np.allclose(eb, ref[1], rtol=0, atol=1e-4) and np.allclose(es, ref[2], rtol=0, atol=1e-4)


# ------------------------------------------------------------------------------------------------ 256 frames per MODCOD
def load256(oracle, modcod):
    import sys

    sys.path.insert(0, GOLD)
    import ldpc_inputs as gi

    return gi.load_fixture(oracle, modcod)


@pytest.mark.parametrize("modcod", ALL)
def test_ldpc_256_frames_across_the_waterfall(oracle, modcod):
    """Oracle == sequential reference decoders (float, int16, int8: bits and CWD on all 256 frames, FER ~1 -> 0), == both
    inter-frame SIMD decoders (ONMS_inter, row a8; templated ..._inter<Update_rule_NMS_simd>, row a9) on every frame
    those report as converged, and the BCH stage behind it."""
    g, d, q16, q8, want = load256(oracle, modcod)
    N, K = MODCODS[modcod]["N_bch"], MODCODS[modcod]["K_bch"]
    thr = os.cpu_count() or 4
    V, cwd, nit = oracle.ldpc_decode_mt(N, d["llr"], thr)
    assert np.array_equal(V, want["seq_f32"]) and np.array_equal(cwd, g["seq_f32_cwd"]) and np.array_equal(nit, g["seq_f32_nit"])
    assert 0 < (cwd == 0).sum() < 256 and nit.min() < 10  # the set really spans failing and early-stopping frames
    for key in ("onms_f32", "tint_f32"):
        ok = g[key + "_cwd"] == 1
        assert ok.sum() >= 64 and np.array_equal(V[ok], want[key][ok]) and (cwd[ok] == 1).all()
    assert np.array_equal(g["onms_f32_cwd"], g["tint_f32_cwd"]) and np.array_equal(want["onms_f32"], want["tint_f32"])
    V16, c16, _ = oracle.ldpc_decode_mt(N, q16, thr, prec="i16")
    assert np.array_equal(V16, want["seq_i16"]) and np.array_equal(c16, g["seq_i16_cwd"])
    for key in ("onms_i16", "tint_i16"):
        ok = g[key + "_cwd"] == 1
        assert np.array_equal(V16[ok], want[key][ok])
    V8, c8, _ = oracle.ldpc_decode_mt(N, q8, thr, prec="i8")
    assert np.array_equal(V8, want["seq_i8"]) and np.array_equal(c8, g["seq_i8_cwd"])
    Vb, cb, _ = oracle.bch_decode(K, N, V)
    assert np.array_equal(Vb, want["bch"]) and np.array_equal(cb, g["bch_cwd"])
