"""GPU parity for the receiver-stage neighbours of the decode chain: Framer::remove_plh (exact copy),
Estimator_DVBS2::estimate (float tolerance: same summation order, libm vs device log10/pow), and the fused PL-frame stage."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, synth_frames

pytestmark = pytest.mark.gpu

ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


@pytest.mark.parametrize("modcod", ALL)
@pytest.mark.parametrize("F", [1, 37])
def test_remove_plh_is_the_reference_copy(oracle, modcod, F):
    from dvbs2_b200 import Framer_DVBS2_cuda, Handle

    n_sym = N_LDPC // MODCODS[modcod]["bps"]
    s = synth_frames(oracle, modcod, F, 6.0, seed=5)
    pl = oracle.insert_plh(n_sym, s["rx"], seed=F)
    h = Handle(modcod=modcod, n_frames=F)
    fr = Framer_DVBS2_cuda(h, F)
    assert fr.pl_frame_size == 2 * oracle.pl_frame_symbols(n_sym) and fr.xfec_frame_size == 2 * n_sym
    out = np.zeros((F, 2 * n_sym), np.float32)
    fr.remove_plh(pl, out)
    assert np.array_equal(out, oracle.remove_plh(n_sym, pl)) and np.array_equal(out, s["rx"])
    h.close()


@pytest.mark.parametrize("modcod,ebn0", [("QPSK-S_8/9", 4.0), ("QPSK-S_3/5", 2.2), ("8PSK-S_3/5", 4.0), ("8PSK-S_8/9", 7.5),
                                         ("16APSK-S_8/9", 7.6), ("QPSK-S_8/9", 40.0)])
def test_estimator_matches_the_reference_formula(oracle, modcod, ebn0):
    from dvbs2_b200 import Estimator_DVBS2_cuda, Handle

    mc = MODCODS[modcod]
    n_sym, F = N_LDPC // mc["bps"], 45
    s = synth_frames(oracle, modcod, F, ebn0, seed=8)
    h = Handle(modcod=modcod, n_frames=F)
    sig, eb, es = (np.zeros(F, np.float32) for _ in range(3))
    Estimator_DVBS2_cuda(h, F).estimate(s["rx"], sig, eb, es)
    so, ebo, eso = oracle.estimate(n_sym, mc["K_bch"], mc["bps"], s["rx"])
    # the moments are summed in the reference's order without contraction (bit-identical); log10f / powf differ from libm
    # by a few ulp, which the cancellation in Ne = |M2 - Se| does not amplify (it happens before them)
    assert np.allclose(es, eso, rtol=0, atol=2e-5 * np.maximum(1, np.abs(eso)).max())
    assert np.allclose(eb, ebo, rtol=0, atol=2e-5 * np.maximum(1, np.abs(ebo)).max())
    assert np.max(np.abs(sig - so) / so) < 5e-6
    h.close()


@pytest.mark.parametrize("modcod,ebn0,packed", [("QPSK-S_8/9", 4.2, False), ("QPSK-S_8/9", 4.2, True), ("8PSK-S_3/5", 4.2, True)])
def test_fused_receiver_stage_from_pl_frames(oracle, modcod, ebn0, packed):
    """PL frames in, descrambled payload out, sigma estimated on the device: equals the chain on the XFEC symbols fed with
    the GPU's own estimates, followed by the oracle's BB descrambler; accepted frames give back the payload."""
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle, Receiver_chain_DVBS2_cuda

    mc = MODCODS[modcod]
    K, n_sym, F = mc["K_bch"], N_LDPC // mc["bps"], 70
    s = synth_frames(oracle, modcod, F, ebn0, seed=21)
    pl = oracle.insert_plh(n_sym, s["rx"], seed=1)
    h = Handle(modcod=modcod, n_frames=F)
    rx = Receiver_chain_DVBS2_cuda(h, F)
    sig = np.zeros(F, np.float32)
    c1, c2, nit = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    V = np.zeros((F, K), np.int32)
    P = np.zeros((F, K // 8), np.uint8)
    rx.receive(pl, SIG=sig, CWD_ldpc=c1, CWD_bch=c2, n_iter=nit, **({"V_K_packed": P} if packed else {"V_K": V}))
    if packed:
        V = np.unpackbits(P, axis=1, bitorder="little").astype(np.int32)
    so, _, _ = oracle.estimate(n_sym, K, mc["bps"], s["rx"])
    assert np.max(np.abs(sig - so) / so) < 5e-6
    # the unfused chain on the same symbols with those sigmas
    Vc = np.zeros((F, K), np.int32)
    d1, d2, dn = np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    Decoder_chain_DVBS2_cuda(h, F).demodulate_decode(sig, s["rx"], d1, d2, Vc, dn)
    assert np.array_equal(c1, d1) and np.array_equal(c2, d2) and np.array_equal(nit, dn)
    assert np.array_equal(V, oracle.bb_scramble(K, Vc))
    good = c2 == 1
    assert good.sum() > F // 2 and np.array_equal(V[good], s["U"][good])
    # device-resident entry point, caller-provided sigma, no descrambling
    dev = torch.device("cuda:0")
    dV = torch.zeros((F, K), dtype=torch.int32, device=dev)
    rx.receive_dev(torch.from_numpy(pl).to(dev), CP=torch.from_numpy(sig).to(dev), V_K=dV, descramble=False)
    h.sync()
    assert np.array_equal(dV.cpu().numpy(), Vc)
    h.close()
