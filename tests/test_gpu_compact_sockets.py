"""GPU parity of the compact socket formats: half / sc16 / sc8 symbols in, int8 LLRs in, packed bytes out.
Bar: bit-identical to the float-socket entry points fed with the same sample values (the device converts exactly as a host
adaptor would), and — for int8 LLRs — to the oracle's / reference's int16 decoder on the same values."""
import numpy as np
import pytest

from oracle.oracle import MODCODS, N_LDPC, synth_frames

pytestmark = pytest.mark.gpu


def _chain_f32(modcod, cp, sym, F):
    from dvbs2_b200 import Decoder_chain_DVBS2_cuda, Handle

    K = MODCODS[modcod]["K_bch"]
    h = Handle(modcod=modcod, n_frames=F)
    V, c1, c2, nit = np.zeros((F, K), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    Decoder_chain_DVBS2_cuda(h, F).demodulate_decode(cp, np.ascontiguousarray(sym), c1, c2, V, nit)
    h.close()
    return V, c1, c2, nit


@pytest.mark.parametrize("modcod,ebn0", [("QPSK-S_8/9", 4.0), ("8PSK-S_3/5", 3.9), ("16APSK-S_8/9", 7.8)])
@pytest.mark.parametrize("fmt,scale", [("sc16", 1.0 / 8192), ("sc8", 1.0 / 40), ("f16", 1.0), ("sc12", 1.0 / 512)])
def test_compact_symbol_formats_equal_float_sockets_on_the_same_samples(oracle, modcod, ebn0, fmt, scale):
    from dvbs2_b200 import Decoder_chain_compact_DVBS2_cuda, Handle
    from dvbs2_b200.modules import iq_pack, iq_unpack

    F, K = 70, MODCODS[modcod]["K_bch"]
    s = synth_frames(oracle, modcod, F, ebn0, seed=31)
    wire = np.ascontiguousarray(iq_pack(s["rx"], fmt, scale))
    assert wire.nbytes * {"sc8": 8, "sc12": 8, "sc16": 4, "f16": 4}[fmt] == s["rx"].nbytes * {"sc8": 2, "sc12": 3, "sc16": 2, "f16": 2}[fmt]
    want = _chain_f32(modcod, s["sigma"], iq_unpack(wire, fmt, scale), F)  # float sockets, same sample values
    h = Handle(modcod=modcod, n_frames=F)
    m = Decoder_chain_compact_DVBS2_cuda(h, F, fmt, scale)
    V, c1, c2, nit = np.zeros((F, K), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    m.demodulate_decode(s["sigma"], wire, c1, c2, V_K=V, n_iter=nit)
    for a, b in zip((V, c1, c2, nit), want):
        assert np.array_equal(a, b)
    packed = np.zeros((F, K // 8), np.uint8)
    m.demodulate_decode(s["sigma"], wire, V_K_packed=packed, descramble=True)
    assert np.array_equal(packed, np.packbits(oracle.bb_scramble(K, V).astype(np.uint8), axis=-1, bitorder="little"))
    with pytest.raises(Exception):
        m.demodulate_decode(s["sigma"], wire)  # no output format
    with pytest.raises(Exception):
        m.demodulate_decode(s["sigma"], wire, V_K=V, V_K_packed=packed)  # two output formats
    h.close()
    if fmt == "sc16":  # 16-bit transport of these symbols loses nothing the decoder notices: same payload as the float run
        ref = _chain_f32(modcod, s["sigma"], s["rx"], F)
        assert (ref[2] == c2).mean() > 0.95


def test_compact_pl_frames_with_on_device_estimator(oracle):
    """sc16 PL frames, sigma estimated on the device: equal to dvbs2_cuda_rx_chain on the same samples as floats"""
    from dvbs2_b200 import Decoder_chain_compact_DVBS2_cuda, Handle, Receiver_chain_DVBS2_cuda
    from dvbs2_b200.modules import iq_pack, iq_unpack

    modcod, F = "8PSK-S_8/9", 40
    mc = MODCODS[modcod]
    K, n_sym = mc["K_bch"], N_LDPC // mc["bps"]
    s = synth_frames(oracle, modcod, F, 7.0, seed=32)
    pl = oracle.insert_plh(n_sym, s["rx"], seed=5)
    wire = np.ascontiguousarray(iq_pack(pl, "sc16", 1.0 / 4096))
    plf = np.ascontiguousarray(iq_unpack(wire, "sc16", 1.0 / 4096))
    h = Handle(modcod=modcod, n_frames=F)
    a = dict(V=np.zeros((F, K // 8), np.uint8), sig=np.zeros(F, np.float32), c2=np.zeros(F, np.int8), nit=np.zeros(F, np.int32))
    b = dict(V=np.zeros((F, K // 8), np.uint8), sig=np.zeros(F, np.float32), c2=np.zeros(F, np.int8), nit=np.zeros(F, np.int32))
    Receiver_chain_DVBS2_cuda(h, F).receive(plf, SIG=a["sig"], CWD_bch=a["c2"], V_K_packed=a["V"], n_iter=a["nit"])
    Decoder_chain_compact_DVBS2_cuda(h, F, "sc16", 1.0 / 4096, pl_frames=True).demodulate_decode(
        None, wire, CWD_bch=b["c2"], V_K_packed=b["V"], n_iter=b["nit"], descramble=True, SIG=b["sig"])
    h.close()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert a["c2"].sum() > F // 2


@pytest.mark.parametrize("modcod,ebn0", [("QPSK-S_8/9", 4.05), ("QPSK-S_3/5", 2.5)])
def test_int8_llr_socket_is_bit_exact_with_the_int16_decoder(oracle, modcod, ebn0):
    from dvbs2_b200 import Decoder_llr8_DVBS2_cuda, Handle

    F, mc = 200, MODCODS[modcod]
    K, N = mc["K_bch"], mc["N_bch"]
    s = synth_frames(oracle, modcod, F, ebn0, seed=33)
    llr = oracle.demodulate(mc["bps"], s["sigma"], s["rx"])
    q8 = oracle.quantize(llr, 2, 6, "i8")  # Quantizer_pow2(6 bits, 2 fractional): |q| <= 31, fits the int8 wire format
    Vl, cl, nl = oracle.ldpc_decode(N, q8.astype(np.int16), prec="i16")
    Vb, cb, _ = oracle.bch_decode(K, N, Vl)
    h = Handle(modcod=modcod, precision="i16", n_frames=F)
    m = Decoder_llr8_DVBS2_cuda(h, F)
    V, c1, c2, nit = np.zeros((F, K), np.int32), np.zeros(F, np.int8), np.zeros(F, np.int8), np.zeros(F, np.int32)
    m.decode(q8, c1, c2, V_K=V, n_iter=nit)
    assert np.array_equal(nit, nl) and np.array_equal(c1, cl) and np.array_equal(c2, cb) and np.array_equal(V, Vb)
    packed = np.zeros((F, K // 8), np.uint8)
    m.decode(q8, V_K_packed=packed, descramble=True)
    assert np.array_equal(packed, np.packbits(oracle.bb_scramble(K, Vb).astype(np.uint8), axis=-1, bitorder="little"))
    st = h.stats()
    hist = np.bincount(np.concatenate([nl, nl]), minlength=33)  # two calls
    assert st["iter_hist"] == [int(x) for x in hist[:33]] and sum(st["iter_hist"]) == st["frames"] == 2 * F
    h.close()
    hf = Handle(modcod=modcod, n_frames=F)
    with pytest.raises(Exception):
        Decoder_llr8_DVBS2_cuda(hf, F).decode(q8, V_K=V)  # float handle: the int8 socket feeds the fixed-point decoder only
    hf.close()


def test_async_entry_point_two_handles_one_host_thread(oracle):
    """dvbs2_cuda_chain_compact_async + dvbs2_cuda_wait: one host thread queues a call on each of two handles (page-locked
    buffers), then waits for both — same bytes as the synchronous entry point."""
    import torch

    from dvbs2_b200 import Decoder_chain_compact_DVBS2_cuda, Handle
    from dvbs2_b200.modules import iq_pack

    modcod, F = "QPSK-S_8/9", 2048 + 70
    K = MODCODS[modcod]["K_bch"]
    s = synth_frames(oracle, modcod, 70, 4.0, seed=5)
    reps = (F + 69) // 70
    rx = np.tile(s["rx"], (reps, 1))[:F]
    sig = np.tile(s["sigma"], reps)[:F]
    wire = np.ascontiguousarray(iq_pack(rx, "sc16", 1.0 / 8192))
    h0 = Handle(modcod=modcod, n_frames=F)
    m0 = Decoder_chain_compact_DVBS2_cuda(h0, F, "sc16", 1.0 / 8192)
    want = np.zeros((F, K // 8), np.uint8)
    want_c = np.zeros(F, np.int8)
    m0.demodulate_decode(sig, wire, CWD_bch=want_c, V_K_packed=want, descramble=True)
    h0.close()
    assert np.array_equal(want[:70], want[70:140])  # tiled input, tiled output

    def pinned(a):
        t = torch.from_numpy(a.copy()).pin_memory()
        return t, t.numpy()

    clones = []
    for _ in range(2):
        h = Handle(modcod=modcod, n_frames=F, host_pipeline_chunks=2)
        m = Decoder_chain_compact_DVBS2_cuda(h, F, "sc16", 1.0 / 8192)
        keep = [pinned(sig), pinned(wire), pinned(np.zeros((F, K // 8), np.uint8)), pinned(np.zeros(F, np.int8))]
        clones.append((h, m, keep))
    for _ in range(2):  # twice: the second round reuses the device buffers of the first
        for h, m, keep in clones:
            keep[2][1][:] = 0
            m.demodulate_decode_async(keep[0][1], keep[1][1], CWD_bch=keep[3][1], V_K_packed=keep[2][1], descramble=True)
        for h, m, keep in clones:
            m.wait()
            assert np.array_equal(keep[2][1], want) and np.array_equal(keep[3][1], want_c)
    for h, _, _ in clones:
        h.close()
