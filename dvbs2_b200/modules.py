"""Host-side mirror of the AFF3CT/StreamPU module surface for the decode chain, on top of the C ABI.

Same task names, argument order and meaning as the reference modules they stand in for (citations relative to the
reference tree):
  Modem_DVBS2_cuda.demodulate(CP, Y_N1, Y_N2)         <- aff3ct::module::Modem::demodulate          (Modem.hxx:222-240)
  Interleaver_DVBS2_cuda.deinterleave(itl, nat)        <- aff3ct::module::Interleaver::deinterleave  (Interleaver.hxx:228-258)
  Decoder_LDPC_DVBS2_cuda.decode_siho(Y_N, CWD, V_K)   <- aff3ct::module::Decoder_SIHO::decode_siho  (Decoder_SIHO.hxx:57-79)
  Decoder_BCH_DVBS2_cuda.decode_hiho(Y_N, CWD, V_K)    <- aff3ct::module::Decoder_HIHO::decode_hiho  (Decoder_HIHO.hxx:42-63)
  Decoder_chain_DVBS2_cuda.demodulate_decode(...)      <- the four tasks bound as in mains/TX_RX_BB/main.cpp:87-91
Like a StreamPU task with set_single_wave(true) (Module.cpp:88-108) every call processes the whole -F batch at once:
buffers are frame-major [n_frames][elements]; a wrong length raises ValueError (the reference throws
spu::tools::length_error), a CUDA failure raises capi.Dvbs2CudaError (the reference throws runtime_error).
The C++ twins that really derive from aff3ct::module::* live in module/ (see INTEGRATION.md).

Host buffers are numpy arrays; the *_dev methods take torch CUDA tensors and run asynchronously on the handle's stream.
Stream-ordering contract of the *_dev methods: the handle launches on its OWN non-blocking stream (Handle.stream()), which
does not synchronise with torch's current stream.  Whatever produced the input tensors (and zero-filled the outputs) on
torch's stream must have finished first — call torch.cuda.synchronize(), or do the torch work under
torch.cuda.ExternalStream(h.stream()) — and results are valid after Handle.sync().
"""
import ctypes as C

import numpy as np

from . import capi

N_LDPC = capi.N_LDPC


def _np_ptr(a, dtype, n_elem, name):
    if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags["C_CONTIGUOUS"]:
        raise ValueError("'%s' must be a C-contiguous numpy array of %s" % (name, np.dtype(dtype).name))
    if a.size != n_elem:
        raise ValueError("'%s.size()' has to be equal to %d ('%s.size()' = %d)" % (name, n_elem, name, a.size))
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t, dtype_name, n_elem, name):
    import torch

    want = {"float32": torch.float32, "float16": torch.float16, "int32": torch.int32, "int16": torch.int16,
            "int8": torch.int8, "uint8": torch.uint8}[dtype_name]
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == want and t.is_contiguous()):
        raise ValueError("'%s' must be a contiguous CUDA tensor of %s" % (name, dtype_name))
    if t.numel() != n_elem:
        raise ValueError("'%s.numel()' has to be equal to %d (is %d)" % (name, n_elem, t.numel()))
    return C.c_void_p(t.data_ptr())


class Handle:
    """One libdvbs2_cuda handle = one device + stream + device buffers (one per module clone, not thread-safe)."""

    def __init__(self, modcod="QPSK-S_8/9", n_ite=10, norm=1.0, precision="f32", early_stop=True, n_frames=1024,
                 device=0, quant_bits=6, quant_dec=2, compaction=True, compaction_rounds=0, compaction_exit_eighths=0,
                 compaction_nf_later=0, ldpc_frames_per_lane=0, ldpc_coop_max_frames=0, host_pipeline_chunks=0, ldpc_onchip=0):
        self.lib = capi.load()
        if modcod not in capi.MODCOD:
            raise ValueError(str(modcod) + " mod-cod scheme not supported.")  # DVBS2.cpp:319
        cfg = capi.Cfg()
        self.lib.dvbs2_cuda_default_cfg(C.byref(cfg))
        cfg.modcod = capi.MODCOD[modcod]
        cfg.n_ite, cfg.norm, cfg.precision = int(n_ite), float(norm), capi.PREC[precision]
        cfg.early_stop, cfg.max_frames, cfg.device = int(bool(early_stop)), int(n_frames), int(device)
        cfg.quant_bits, cfg.quant_dec = int(quant_bits), int(quant_dec)
        cfg.no_compaction = 0 if compaction else 1
        cfg.compaction_rounds, cfg.compaction_exit_eighths = int(compaction_rounds), int(compaction_exit_eighths)
        cfg.compaction_nf_later = int(compaction_nf_later)
        cfg.ldpc_frames_per_lane = int(ldpc_frames_per_lane)
        cfg.ldpc_coop_max_frames = int(ldpc_coop_max_frames)
        cfg.host_pipeline_chunks = int(host_pipeline_chunks)
        cfg.ldpc_onchip = int(ldpc_onchip)
        self.modcod, self.precision, self.device = modcod, precision, device
        self._h = C.c_void_p()
        rc = self.lib.dvbs2_cuda_create(C.byref(cfg), C.byref(self._h))
        if rc != capi.OK:
            self._h = C.c_void_p()
            raise capi.Dvbs2CudaError(rc, capi.strerror(rc))
        info = capi.Info()
        self.lib.dvbs2_cuda_get_info(self._h, C.byref(info))
        self.info = {n: getattr(info, n) for n, _ in capi.Info._fields_}
        self.llr_dtype = {"f32": np.float32, "i16": np.int16, "i8": np.int8}[precision]

    def check(self, rc):
        if rc != capi.OK:
            msg = capi.strerror(rc)
            if rc == capi.ERR_CUDA:
                msg += ": " + self.lib.dvbs2_cuda_last_error(self._h).decode()
            raise capi.Dvbs2CudaError(rc, msg)

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            self.lib.dvbs2_cuda_destroy(self._h)
            self._h = C.c_void_p()

    __del__ = close

    def sync(self):
        self.check(self.lib.dvbs2_cuda_sync(self._h))

    def stream(self):
        return self.lib.dvbs2_cuda_stream(self._h)

    def stats(self):
        st = capi.Stats()
        self.check(self.lib.dvbs2_cuda_get_stats(self._h, C.byref(st)))
        out = {n: int(getattr(st, n)) for n, _ in capi.Stats._fields_ if n != "iter_hist"}
        out["iter_hist"] = [int(x) for x in st.iter_hist]  # frames that left the LDPC decoder after i iterations
        return out

    def reset_stats(self):
        self.check(self.lib.dvbs2_cuda_reset_stats(self._h))

    def set_profiling(self, enable=True):
        """Per-stage device timers (CUDA events on the launching stream around every kernel of the chain)."""
        self.check(self.lib.dvbs2_cuda_set_profiling(self._h, int(bool(enable))))

    def timers(self):
        t = capi.Timers()
        self.check(self.lib.dvbs2_cuda_get_timers(self._h, C.byref(t)))
        return {n: getattr(t, n) for n, _ in capi.Timers._fields_}

    def host_register(self, arr):
        self.check(self.lib.dvbs2_cuda_host_register(self._h, arr.ctypes.data_as(C.c_void_p), arr.nbytes))

    def host_unregister(self, arr):
        self.check(self.lib.dvbs2_cuda_host_unregister(self._h, arr.ctypes.data_as(C.c_void_p)))


class _Module:
    """Common part: a (possibly shared) Handle and the StreamPU-style n_frames bookkeeping."""

    def __init__(self, handle, n_frames):
        self.h = handle
        self.n_frames = 0
        self.set_n_frames(n_frames)

    def set_n_frames(self, n_frames):  # Module::set_n_frames (Module.cpp:43-86)
        if n_frames <= 0:
            raise ValueError("'n_frames' has to be greater than 0 ('n_frames' = %d)." % n_frames)
        if n_frames > self.h.info["max_frames"]:
            raise ValueError("'n_frames' exceeds the handle's max_frames (%d > %d)" % (n_frames, self.h.info["max_frames"]))
        self.n_frames = int(n_frames)

    def get_n_frames(self):
        return self.n_frames


class Modem_DVBS2_cuda(_Module):
    """Soft demapper of the handle's MODCOD.  demodulate: CP[F] (sigma), Y_N1[F][2*n_sym] -> Y_N2[F][16200]."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.N = N_LDPC
        self.N_fil = 2 * handle.info["n_sym"]

    def demodulate(self, CP, Y_N1, Y_N2):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_demodulate(h._h, _np_ptr(CP, np.float32, F, "CP"),
                                            _np_ptr(Y_N1, np.float32, F * self.N_fil, "Y_N1"),
                                            _np_ptr(Y_N2, np.float32, F * self.N, "Y_N2"), F))

    def demodulate_dev(self, CP, Y_N1, Y_N2, deinterleave=False):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_demodulate_dev(h._h, _dev_ptr(CP, "float32", F, "CP"),
                                                _dev_ptr(Y_N1, "float32", F * self.N_fil, "Y_N1"),
                                                _dev_ptr(Y_N2, "float32", F * self.N, "Y_N2"), int(deinterleave), F))


class Interleaver_DVBS2_cuda(_Module):
    """Column-row deinterleaver of the handle's MODCOD (identity for QPSK)."""

    def deinterleave(self, itl, nat):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_deinterleave(h._h, _np_ptr(itl, np.float32, F * N_LDPC, "itl"),
                                              _np_ptr(nat, np.float32, F * N_LDPC, "nat"), F))


class Decoder_LDPC_DVBS2_cuda(_Module):
    """Horizontal-layered (N)MS decoder.  decode_siho: Y_N[F][16200] -> CWD[F] (int8), V_K[F][K] (int32)."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K, self.N = handle.info["K_ldpc"], N_LDPC

    def decode_siho(self, Y_N, CWD, V_K, n_iter=None):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_ldpc_decode_siho(
            h._h, _np_ptr(Y_N, h.llr_dtype, F * self.N, "Y_N"), _np_ptr(CWD, np.int8, F, "CWD"),
            _np_ptr(V_K, np.int32, F * self.K, "V_K"),
            None if n_iter is None else _np_ptr(n_iter, np.int32, F, "n_iter"), F))
        return int(np.count_nonzero(CWD == 0) != 0)  # task status: 0 = every frame is a codeword

    def decode_siho_dev(self, Y_N, CWD, V_K, n_iter=None):
        F, h = self.n_frames, self.h
        name = {"f32": "float32", "i16": "int16", "i8": "int8"}[h.precision]
        h.check(h.lib.dvbs2_cuda_ldpc_decode_siho_dev(
            h._h, _dev_ptr(Y_N, name, F * self.N, "Y_N"), _dev_ptr(CWD, "int8", F, "CWD"),
            _dev_ptr(V_K, "int32", F * self.K, "V_K"),
            None if n_iter is None else _dev_ptr(n_iter, "int32", F, "n_iter"), F))


class Decoder_BCH_DVBS2_cuda(_Module):
    """BCH t = 12 hard decoder, DVB-S2 bit order.  decode_hiho: Y_N[F][N_bch] -> CWD[F], V_K[F][K_bch]."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K, self.N = handle.info["K_bch"], handle.info["N_bch"]

    def decode_hiho(self, Y_N, CWD, V_K):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_bch_decode_hiho(h._h, _np_ptr(Y_N, np.int32, F * self.N, "Y_N"),
                                                 _np_ptr(CWD, np.int8, F, "CWD"),
                                                 _np_ptr(V_K, np.int32, F * self.K, "V_K"), F))
        return int(np.count_nonzero(CWD == 0) != 0)

    def decode_hiho_dev(self, Y_N, CWD, V_K):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_bch_decode_hiho_dev(h._h, _dev_ptr(Y_N, "int32", F * self.N, "Y_N"),
                                                     _dev_ptr(CWD, "int8", F, "CWD"),
                                                     _dev_ptr(V_K, "int32", F * self.K, "V_K"), F))


class Decoder_chain_DVBS2_cuda(_Module):
    """Fused stage: demodulate -> deinterleave -> (quantize) -> LDPC decode_siho -> BCH decode_hiho on the device."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K = handle.info["K_bch"]
        self.N_fil = 2 * handle.info["n_sym"]

    def demodulate_decode(self, CP, Y_N1, CWD_ldpc, CWD_bch, V_K, n_iter=None):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_chain(
            h._h, _np_ptr(CP, np.float32, F, "CP"), _np_ptr(Y_N1, np.float32, F * self.N_fil, "Y_N1"),
            None if CWD_ldpc is None else _np_ptr(CWD_ldpc, np.int8, F, "CWD_ldpc"),
            None if CWD_bch is None else _np_ptr(CWD_bch, np.int8, F, "CWD_bch"),
            _np_ptr(V_K, np.int32, F * self.K, "V_K"),
            None if n_iter is None else _np_ptr(n_iter, np.int32, F, "n_iter"), F))

    def demodulate_decode_dev(self, CP, Y_N1, CWD_ldpc, CWD_bch, V_K, n_iter=None):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_chain_dev(
            h._h, _dev_ptr(CP, "float32", F, "CP"), _dev_ptr(Y_N1, "float32", F * self.N_fil, "Y_N1"),
            None if CWD_ldpc is None else _dev_ptr(CWD_ldpc, "int8", F, "CWD_ldpc"),
            None if CWD_bch is None else _dev_ptr(CWD_bch, "int8", F, "CWD_bch"),
            _dev_ptr(V_K, "int32", F * self.K, "V_K"),
            None if n_iter is None else _dev_ptr(n_iter, "int32", F, "n_iter"), F))


    def demodulate_decode_packed(self, CP, Y_N1, CWD_ldpc, CWD_bch, V_K_packed, n_iter=None, descramble=False):
        """Same stage, bit-packed output [F][K_bch/8] uint8 (Bit_packer layout, LSB first), optionally BB-descrambled."""
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_chain_packed(
            h._h, _np_ptr(CP, np.float32, F, "CP"), _np_ptr(Y_N1, np.float32, F * self.N_fil, "Y_N1"),
            None if CWD_ldpc is None else _np_ptr(CWD_ldpc, np.int8, F, "CWD_ldpc"),
            None if CWD_bch is None else _np_ptr(CWD_bch, np.int8, F, "CWD_bch"),
            _np_ptr(V_K_packed, np.uint8, F * (self.K // 8), "V_K_packed"),
            None if n_iter is None else _np_ptr(n_iter, np.int32, F, "n_iter"), int(bool(descramble)), F))

    def demodulate_decode_packed_dev(self, CP, Y_N1, CWD_ldpc, CWD_bch, V_K_packed, n_iter=None, descramble=False):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_chain_packed_dev(
            h._h, _dev_ptr(CP, "float32", F, "CP"), _dev_ptr(Y_N1, "float32", F * self.N_fil, "Y_N1"),
            None if CWD_ldpc is None else _dev_ptr(CWD_ldpc, "int8", F, "CWD_ldpc"),
            None if CWD_bch is None else _dev_ptr(CWD_bch, "int8", F, "CWD_bch"),
            _dev_ptr(V_K_packed, "uint8", F * (self.K // 8), "V_K_packed"),
            None if n_iter is None else _dev_ptr(n_iter, "int32", F, "n_iter"), int(bool(descramble)), F))


IQ_DTYPE = {"f32": np.float32, "f16": np.float16, "sc16": np.int16, "sc8": np.int8, "sc12": np.uint8}  # sc12: 3 bytes per symbol


def iq_pack(Y_N1, iq_format, iq_scale=1.0):
    """float I/Q -> wire format, what the upstream adaptor (module/Adaptor_DVBS2_cuda.hpp, task `pack`) does on the host:
    f16 = IEEE round-to-nearest; sc16 / sc8 = round-half-away(x / scale) saturated to the type (a radio's ADC word)."""
    Y_N1 = np.asarray(Y_N1, np.float32)
    if iq_format == "f32":
        return Y_N1
    if iq_format == "f16":
        return Y_N1.astype(np.float16)
    lim = {"sc16": 32767, "sc8": 127, "sc12": 2047}[iq_format]
    q = Y_N1 * (np.float32(1.0) / np.float32(iq_scale))  # same two float operations as IQ_packer<Q> (Adaptor_DVBS2_cuda.hpp)
    q = np.sign(q) * np.floor(np.abs(q) + np.float32(0.5))
    if iq_format == "sc12":  # [..., 2 n] components -> [..., 3 n] bytes: I = bits 0..11, Q = bits 12..23 of a little-endian word
        v = np.clip(q, -lim, lim).astype(np.int32) & 0xfff
        w = v[..., 0::2] | (v[..., 1::2] << 12)
        return np.stack([w & 0xff, (w >> 8) & 0xff, (w >> 16) & 0xff], axis=-1).astype(np.uint8).reshape(*w.shape[:-1], -1)
    return np.clip(q, -lim, lim).astype(IQ_DTYPE[iq_format])


def iq_unpack(Y, iq_format, iq_scale=1.0):
    """wire format -> the float samples the device computes with ((float)x * scale, one rounded multiply)"""
    if iq_format in ("f32", "f16"):
        return np.asarray(Y).astype(np.float32)
    if iq_format == "sc12":
        b = np.asarray(Y, np.uint8).astype(np.int32)
        b = b.reshape(*b.shape[:-1], -1, 3)
        w = b[..., 0] | (b[..., 1] << 8) | (b[..., 2] << 16)
        comp = np.stack([w & 0xfff, (w >> 12) & 0xfff], axis=-1).reshape(*w.shape[:-1], -1)
        comp = (comp ^ 0x800) - 0x800  # sign-extend 12 bits
        return comp.astype(np.float32) * np.float32(iq_scale)
    return np.asarray(Y).astype(np.float32) * np.float32(iq_scale)


class Decoder_chain_compact_DVBS2_cuda(_Module):
    """The fused stage behind compact sockets: symbols as f16 / sc16 / sc8 instead of float (XFEC frames, or
    whole PL frames with pl_frames=True), payload as packed bytes instead of int32 per bit."""

    def __init__(self, handle, n_frames=1, iq_format="sc16", iq_scale=1.0 / 8192, pl_frames=False):
        super().__init__(handle, n_frames)
        if iq_format not in capi.IQ_FORMAT:
            raise ValueError("'iq_format' has to be one of %s" % sorted(capi.IQ_FORMAT))
        self.K = handle.info["K_bch"]
        self.iq_format, self.iq_scale, self.pl_frames = iq_format, float(iq_scale), bool(pl_frames)
        self.N_fil = (handle.lib.dvbs2_cuda_pl_frame_size(handle._h) if pl_frames else 2 * handle.info["n_sym"])
        if iq_format == "sc12":
            self.N_fil = self.N_fil // 2 * 3  # bytes per frame: the socket is a byte array
        assert handle.lib.dvbs2_cuda_iq_frame_bytes(handle._h, capi.IQ_FORMAT[iq_format], int(pl_frames)) == \
            self.N_fil * np.dtype(IQ_DTYPE[iq_format]).itemsize

    def demodulate_decode(self, CP, Y_N1, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None, descramble=False,
                          SIG=None):
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, name: None if a is None else _np_ptr(a, dt, n, name)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_chain_compact(
            h._h, capi.IQ_FORMAT[self.iq_format], self.iq_scale, opt(CP, np.float32, F, "CP"),
            _np_ptr(Y_N1, IQ_DTYPE[self.iq_format], F * self.N_fil, "Y_N1"), opt(SIG, np.float32, F, "SIG"),
            opt(CWD_ldpc, np.int8, F, "CWD_ldpc"), opt(CWD_bch, np.int8, F, "CWD_bch"), opt(V_K, np.int32, F * self.K, "V_K"),
            opt(V_K_packed, np.uint8, F * (self.K // 8), "V_K_packed"), int(bool(descramble)),
            opt(n_iter, np.int32, F, "n_iter"), int(self.pl_frames), F))

    def demodulate_decode_async(self, CP, Y_N1, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None,
                                descramble=False, SIG=None):
        """Queues the call and returns (dvbs2_cuda_chain_compact_async); wait() blocks until the results are in the host
        buffers.  The buffers must be page-locked and stay untouched until then; one call in flight per handle."""
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, name: None if a is None else _np_ptr(a, dt, n, name)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_chain_compact_async(
            h._h, capi.IQ_FORMAT[self.iq_format], self.iq_scale, opt(CP, np.float32, F, "CP"),
            _np_ptr(Y_N1, IQ_DTYPE[self.iq_format], F * self.N_fil, "Y_N1"), opt(SIG, np.float32, F, "SIG"),
            opt(CWD_ldpc, np.int8, F, "CWD_ldpc"), opt(CWD_bch, np.int8, F, "CWD_bch"), opt(V_K, np.int32, F * self.K, "V_K"),
            opt(V_K_packed, np.uint8, F * (self.K // 8), "V_K_packed"), int(bool(descramble)),
            opt(n_iter, np.int32, F, "n_iter"), int(self.pl_frames), F))

    def wait(self):
        self.h.check(self.h.lib.dvbs2_cuda_wait(self.h._h))

    def demodulate_decode_dev(self, CP, Y_N1, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None,
                              descramble=False, SIG=None):
        F, h = self.n_frames, self.h
        name = {"f32": "float32", "f16": "float16", "sc16": "int16", "sc8": "int8"}[self.iq_format]
        opt = lambda a, dt, n, nm: None if a is None else _dev_ptr(a, dt, n, nm)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_chain_compact_dev(
            h._h, capi.IQ_FORMAT[self.iq_format], self.iq_scale, opt(CP, "float32", F, "CP"),
            _dev_ptr(Y_N1, name, F * self.N_fil, "Y_N1"), opt(SIG, "float32", F, "SIG"), opt(CWD_ldpc, "int8", F, "CWD_ldpc"),
            opt(CWD_bch, "int8", F, "CWD_bch"), opt(V_K, "int32", F * self.K, "V_K"),
            opt(V_K_packed, "uint8", F * (self.K // 8), "V_K_packed"), int(bool(descramble)),
            opt(n_iter, "int32", F, "n_iter"), int(self.pl_frames), F))


class Decoder_llr8_DVBS2_cuda(_Module):
    """Fixed-point decode stage from int8 LLRs (Quantizer_pow2 outputs): int16 LDPC decode_siho -> BCH decode_hiho
    (-> BB descramble), int32 bits or packed bytes out.  The handle must use precision "i16"."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K, self.N = handle.info["K_bch"], N_LDPC

    def decode(self, Y_N, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None, descramble=False):
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, name: None if a is None else _np_ptr(a, dt, n, name)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_decode_llr8(
            h._h, _np_ptr(Y_N, np.int8, F * self.N, "Y_N"), opt(CWD_ldpc, np.int8, F, "CWD_ldpc"),
            opt(CWD_bch, np.int8, F, "CWD_bch"), opt(V_K, np.int32, F * self.K, "V_K"),
            opt(V_K_packed, np.uint8, F * (self.K // 8), "V_K_packed"), int(bool(descramble)),
            opt(n_iter, np.int32, F, "n_iter"), F))

    def decode_dev(self, Y_N, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None, descramble=False):
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, nm: None if a is None else _dev_ptr(a, dt, n, nm)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_decode_llr8_dev(
            h._h, _dev_ptr(Y_N, "int8", F * self.N, "Y_N"), opt(CWD_ldpc, "int8", F, "CWD_ldpc"),
            opt(CWD_bch, "int8", F, "CWD_bch"), opt(V_K, "int32", F * self.K, "V_K"),
            opt(V_K_packed, "uint8", F * (self.K // 8), "V_K_packed"), int(bool(descramble)),
            opt(n_iter, "int32", F, "n_iter"), F))


class Framer_DVBS2_cuda(_Module):
    """Framer<float>::remove_plh on the device (src/common/Module/Framer/Framer.hxx:319-343)."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.xfec_frame_size = 2 * handle.info["n_sym"]
        self.pl_frame_size = handle.lib.dvbs2_cuda_pl_frame_size(handle._h)

    def remove_plh(self, Y_N1, Y_N2):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_remove_plh(h._h, _np_ptr(Y_N1, np.float32, F * self.pl_frame_size, "Y_N1"),
                                            _np_ptr(Y_N2, np.float32, F * self.xfec_frame_size, "Y_N2"), F))


class Estimator_DVBS2_cuda(_Module):
    """Estimator_DVBS2<float>::estimate on the device (src/common/Module/Estimator/Estimator_DVBS2.hxx:31-58)."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.N = 2 * handle.info["n_sym"]

    def estimate(self, X_N, SIG, Eb_N0=None, Es_N0=None):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_estimate(
            h._h, _np_ptr(X_N, np.float32, F * self.N, "X_N"), _np_ptr(SIG, np.float32, F, "SIG"),
            None if Eb_N0 is None else _np_ptr(Eb_N0, np.float32, F, "Eb_N0"),
            None if Es_N0 is None else _np_ptr(Es_N0, np.float32, F, "Es_N0"), F))


    def estimate_dev(self, X_N, SIG, Eb_N0=None, Es_N0=None):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_estimate_dev(
            h._h, _dev_ptr(X_N, "float32", F * self.N, "X_N"), _dev_ptr(SIG, "float32", F, "SIG"),
            None if Eb_N0 is None else _dev_ptr(Eb_N0, "float32", F, "Eb_N0"),
            None if Es_N0 is None else _dev_ptr(Es_N0, "float32", F, "Es_N0"), F))


class Receiver_chain_DVBS2_cuda(_Module):
    """Fused receiver stage from PL frames: remove_plh -> estimate (or caller's sigma) -> demodulate -> deinterleave -> LDPC ->
    BCH -> optional BB descramble; int32 or bit-packed output (SURVEY §8f N1 + N2)."""

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K = handle.info["K_bch"]
        self.pl_frame_size = handle.lib.dvbs2_cuda_pl_frame_size(handle._h)

    def receive(self, PL_frames, CP=None, SIG=None, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None,
                descramble=True):
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, name: None if a is None else _np_ptr(a, dt, n, name)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_rx_chain(
            h._h, _np_ptr(PL_frames, np.float32, F * self.pl_frame_size, "PL_frames"), opt(CP, np.float32, F, "CP"),
            opt(SIG, np.float32, F, "SIG"), opt(CWD_ldpc, np.int8, F, "CWD_ldpc"), opt(CWD_bch, np.int8, F, "CWD_bch"),
            opt(V_K, np.int32, F * self.K, "V_K"), opt(V_K_packed, np.uint8, F * (self.K // 8), "V_K_packed"),
            int(bool(descramble)), opt(n_iter, np.int32, F, "n_iter"), F))

    def receive_dev(self, PL_frames, CP=None, SIG=None, CWD_ldpc=None, CWD_bch=None, V_K=None, V_K_packed=None, n_iter=None,
                    descramble=True):
        F, h = self.n_frames, self.h
        opt = lambda a, dt, n, name: None if a is None else _dev_ptr(a, dt, n, name)  # noqa: E731
        h.check(h.lib.dvbs2_cuda_rx_chain_dev(
            h._h, _dev_ptr(PL_frames, "float32", F * self.pl_frame_size, "PL_frames"), opt(CP, "float32", F, "CP"),
            opt(SIG, "float32", F, "SIG"), opt(CWD_ldpc, "int8", F, "CWD_ldpc"), opt(CWD_bch, "int8", F, "CWD_bch"),
            opt(V_K, "int32", F * self.K, "V_K"), opt(V_K_packed, "uint8", F * (self.K // 8), "V_K_packed"),
            int(bool(descramble)), opt(n_iter, "int32", F, "n_iter"), F))


class Transmitter_DVBS2_cuda(_Module):
    """On-device TX chain + AWGN channel + BFER monitor around the decode path (Monte-Carlo / benchmark harness):
    Source -> Scrambler_BB.scramble -> Encoder_BCH_DVBS2 -> Encoder_LDPC_DVBS2 -> interleave -> modulate -> add_noise
    (mains/TX_RX_BB/main.cpp:75-84) and, behind the chain, Scrambler_BB.descramble -> Monitor_BFER.check_errors (:92-94)."""

    STAGES = {"payload": 0, "scrambled": 1, "bch": 2, "ldpc": 3}

    def __init__(self, handle, n_frames=1):
        super().__init__(handle, n_frames)
        self.K = handle.info["K_bch"]
        self.N_fil = 2 * handle.info["n_sym"]

    def set_payload(self, U):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_tx_set_payload(h._h, _np_ptr(U, np.int32, F * self.K, "U"), F))

    def random_payload(self, seed):
        h = self.h
        h.check(h.lib.dvbs2_cuda_tx_random_payload(h._h, int(seed), self.n_frames))

    def modulate_awgn_dev(self, sigma, seed, CP, Y_N1):
        F, h = self.n_frames, self.h
        h.check(h.lib.dvbs2_cuda_tx_modulate_awgn_dev(h._h, float(sigma), int(seed), _dev_ptr(CP, "float32", F, "CP"),
                                                      _dev_ptr(Y_N1, "float32", F * self.N_fil, "Y_N1"), F))

    def get_bits(self, stage):
        F, h = self.n_frames, self.h
        n = {0: self.K, 1: self.K, 2: h.info["N_bch"], 3: N_LDPC}[self.STAGES[stage]]
        out = np.zeros((F, n), np.int32)
        h.check(h.lib.dvbs2_cuda_tx_get_bits(h._h, self.STAGES[stage], out.ctypes.data_as(C.c_void_p), F))
        return out

    def check_errors(self):
        """Queue the comparison of the last chain output on this handle with the payload (asynchronous)."""
        h = self.h
        h.check(h.lib.dvbs2_cuda_monitor_check(h._h, self.n_frames))

    def monitor(self):
        m = capi.Monitor()
        h = self.h
        h.check(h.lib.dvbs2_cuda_monitor_get(h._h, C.byref(m)))
        return {"n_frames": int(m.n_frames), "n_be": int(m.n_bit_errors), "n_fe": int(m.n_frame_errors)}
