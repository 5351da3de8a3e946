"""GPU parity: on-device transmitter (scrambler, BCH/LDPC encoders, interleaver, mapper), AWGN and BFER monitor vs the oracle."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, ebn0_to_sigma

pytestmark = pytest.mark.gpu

ALL = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


@pytest.mark.parametrize("modcod", ALL)
@pytest.mark.parametrize("F", [70])
def test_tx_bits_and_symbols_bit_exact(oracle, modcod, F):
    """Every TX stage equals the oracle's (which is pinned against the reference's encoders/modem)."""
    import torch

    from dvbs2_b200 import Handle, Transmitter_DVBS2_cuda

    mc = MODCODS[modcod]
    rng = np.random.default_rng(7)
    U = rng.integers(0, 2, size=(F, mc["K_bch"]), dtype=np.int32)
    want = oracle.tx(modcod, U)
    h = Handle(modcod=modcod, n_frames=F)
    tx = Transmitter_DVBS2_cuda(h, F)
    tx.set_payload(U)
    dev = torch.device("cuda:0")
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    sym = torch.zeros((F, 2 * N_LDPC // mc["bps"]), dtype=torch.float32, device=dev)
    tx.modulate_awgn_dev(0.0, 1, cp, sym)
    h.sync()
    assert np.array_equal(tx.get_bits("payload"), U)
    assert np.array_equal(tx.get_bits("scrambled"), want["scr"])
    assert np.array_equal(tx.get_bits("bch"), want["bch"])
    assert np.array_equal(tx.get_bits("ldpc"), want["ldpc"])
    assert np.array_equal(sym.cpu().numpy(), want["sym"])  # noiseless symbols: the constellation floats themselves
    assert (cp.cpu().numpy() == 0).all()
    h.close()


def test_awgn_statistics_and_reproducibility():
    import torch

    from dvbs2_b200 import Handle, Transmitter_DVBS2_cuda

    modcod, F = "QPSK-S_8/9", 256
    h = Handle(modcod=modcod, n_frames=F)
    tx = Transmitter_DVBS2_cuda(h, F)
    tx.random_payload(3)
    dev = torch.device("cuda:0")
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    clean = torch.zeros((F, 16200), dtype=torch.float32, device=dev)
    a = torch.zeros_like(clean)
    b = torch.zeros_like(clean)
    c = torch.zeros_like(clean)
    sigma = 0.3366
    tx.modulate_awgn_dev(0.0, 0, cp, clean)
    tx.modulate_awgn_dev(sigma, 11, cp, a)
    tx.modulate_awgn_dev(sigma, 11, cp, b)
    tx.modulate_awgn_dev(sigma, 12, cp, c)
    h.sync()
    assert torch.equal(a, b) and not torch.equal(a, c)
    n = (a - clean).double()
    assert abs(n.mean().item()) < 2e-3 and abs(n.std().item() / sigma - 1) < 5e-3
    # I and Q and neighbouring symbols uncorrelated, fourth moment Gaussian
    assert abs((n[:, 0::2] * n[:, 1::2]).mean().item()) / sigma**2 < 5e-3
    assert abs((n[:, :-2] * n[:, 2:]).mean().item()) / sigma**2 < 5e-3
    assert abs((n**4).mean().item() / sigma**4 - 3) < 0.05
    assert np.allclose(cp.cpu().numpy(), sigma)
    # payload bits are balanced
    bits = tx.get_bits("payload")
    assert abs(bits.mean() - 0.5) < 5e-3
    h.close()


@pytest.mark.parametrize("modcod,ebn0,prec", [("QPSK-S_8/9", 3.9, "f32"), ("8PSK-S_3/5", 3.9, "f32"), ("16APSK-S_8/9", 7.6, "f32"),
                                              ("8PSK-S_3/5", 3.9, "i16")])
def test_monte_carlo_loop_monitor_matches_oracle(oracle, modcod, ebn0, prec):
    """TX -> AWGN -> fused chain -> monitor entirely on the device; the monitor's BE/FE equal what the oracle chain counts
    on the same received symbols, and BB-descrambled accepted frames equal the payload."""
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle, Modem_DVBS2_cuda, Transmitter_DVBS2_cuda

    mc = MODCODS[modcod]
    F = 96
    h = Handle(modcod=modcod, n_frames=F, precision=prec)
    tx = Transmitter_DVBS2_cuda(h, F)
    ch = Decoder_chain_DVBS2_cuda(h, F)
    dev = torch.device("cuda:0")
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    sym = torch.zeros((F, 2 * N_LDPC // mc["bps"]), dtype=torch.float32, device=dev)
    V = torch.zeros((F, mc["K_bch"]), dtype=torch.int32, device=dev)
    c1 = torch.zeros(F, dtype=torch.int8, device=dev)
    c2 = torch.zeros(F, dtype=torch.int8, device=dev)
    h.reset_stats()
    tx.random_payload(21)
    tx.modulate_awgn_dev(ebn0_to_sigma(modcod, ebn0), 5, cp, sym)
    ch.demodulate_decode_dev(cp, sym, c1, c2, V, None)
    tx.check_errors()
    mon = tx.monitor()
    U = tx.get_bits("payload")
    # oracle on the GPU's LLRs of the same symbols
    llr = np.zeros((F, N_LDPC), np.float32)
    Modem_DVBS2_cuda(h, F).demodulate(cp.cpu().numpy(), sym.cpu().numpy(), llr)
    _, inv = oracle.itl_lut(mc["bps"], mc["itl"])
    nat = np.ascontiguousarray(llr[:, inv])
    if prec != "f32":
        nat = oracle.quantize(nat, 2, 6, prec)
    Vl, cl, _ = oracle.ldpc_decode(mc["N_bch"], nat, prec=prec)
    Vb, cb, _ = oracle.bch_decode(mc["K_bch"], mc["N_bch"], Vl)
    assert np.array_equal(V.cpu().numpy(), Vb) and np.array_equal(c2.cpu().numpy(), cb)
    err = oracle.bb_scramble(mc["K_bch"], Vb) != U
    assert mon == {"n_frames": F, "n_be": int(err.sum()), "n_fe": int(err.any(axis=1).sum())}
    assert 0 < mon["n_fe"] < F  # the operating points are inside the waterfall
    # a second batch accumulates; reset_stats clears
    tx.check_errors()
    assert tx.monitor()["n_frames"] == 2 * F
    h.reset_stats()
    assert tx.monitor() == {"n_frames": 0, "n_be": 0, "n_fe": 0}
    h.close()


def test_monitor_ragged_batch(oracle):
    import torch

    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle, Transmitter_DVBS2_cuda

    modcod, F = "QPSK-S_8/9", 45
    mc = MODCODS[modcod]
    h = Handle(modcod=modcod, n_frames=F)
    tx, ch = Transmitter_DVBS2_cuda(h, F), Decoder_chain_DVBS2_cuda(h, F)
    dev = torch.device("cuda:0")
    cp = torch.zeros(F, dtype=torch.float32, device=dev)
    sym = torch.zeros((F, 16200), dtype=torch.float32, device=dev)
    V = torch.zeros((F, mc["K_bch"]), dtype=torch.int32, device=dev)
    tx.random_payload(2)
    tx.modulate_awgn_dev(ebn0_to_sigma(modcod, 6.0), 9, cp, sym)
    ch.demodulate_decode_dev(cp, sym, None, None, V, None)
    tx.check_errors()
    assert tx.monitor() == {"n_frames": F, "n_be": 0, "n_fe": 0}
    assert np.array_equal(oracle.bb_scramble(mc["K_bch"], V.cpu().numpy()), tx.get_bits("payload"))
    h.close()
