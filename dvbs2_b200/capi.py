"""ctypes binding of libdvbs2_cuda.so (include/dvbs2_cuda.h).

The library is the product; this module only loads it and declares prototypes.  There is no fallback: if the shared
library is missing or no CUDA device is usable, importing works but `load()` / `Handle(...)` raise.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DVBS2_CUDA_LIB") or os.path.join(HERE, "libdvbs2_cuda.so")  # override: kernel-variant experiments

N_LDPC = 16200
OK, ERR_INVALID, ERR_CUDA, ERR_TOO_MANY, ERR_NOMEM = 0, -1, -2, -3, -4
PREC = {"f32": 0, "i16": 1, "i8": 2}
IQ_FORMAT = {"f32": 0, "f16": 1, "sc16": 2, "sc8": 3, "sc12": 4}  # dvbs2_iq_format
MODCOD = {"QPSK-S_8/9": 0, "QPSK-S_3/5": 1, "8PSK-S_3/5": 2, "8PSK-S_8/9": 3, "16APSK-S_8/9": 4}

# every symbol include/dvbs2_cuda.h declares (tests check the library exports all of them)
ABI_VERSION = 2  # DVBS2_CUDA_ABI_VERSION of include/dvbs2_cuda.h

SYMBOLS = [
    "dvbs2_cuda_default_cfg", "dvbs2_cuda_create", "dvbs2_cuda_destroy", "dvbs2_cuda_get_info", "dvbs2_cuda_get_stats",
    "dvbs2_cuda_reset_stats", "dvbs2_cuda_strerror", "dvbs2_cuda_last_error", "dvbs2_cuda_device_count",
    "dvbs2_cuda_demodulate", "dvbs2_cuda_deinterleave", "dvbs2_cuda_ldpc_decode_siho", "dvbs2_cuda_bch_decode_hiho",
    "dvbs2_cuda_quantize", "dvbs2_cuda_chain", "dvbs2_cuda_host_register", "dvbs2_cuda_host_unregister",
    "dvbs2_cuda_stream", "dvbs2_cuda_sync", "dvbs2_cuda_demodulate_dev", "dvbs2_cuda_ldpc_decode_siho_dev",
    "dvbs2_cuda_bch_decode_hiho_dev", "dvbs2_cuda_chain_dev", "dvbs2_cuda_set_profiling", "dvbs2_cuda_get_timers",
    "dvbs2_cuda_tx_set_payload", "dvbs2_cuda_tx_random_payload", "dvbs2_cuda_tx_modulate_awgn_dev", "dvbs2_cuda_tx_get_bits",
    "dvbs2_cuda_monitor_check", "dvbs2_cuda_monitor_get", "dvbs2_cuda_chain_packed", "dvbs2_cuda_chain_packed_dev",
    "dvbs2_cuda_pl_frame_size", "dvbs2_cuda_remove_plh", "dvbs2_cuda_remove_plh_dev", "dvbs2_cuda_estimate",
    "dvbs2_cuda_estimate_dev", "dvbs2_cuda_rx_chain", "dvbs2_cuda_rx_chain_dev", "dvbs2_cuda_iq_frame_bytes",
    "dvbs2_cuda_chain_compact", "dvbs2_cuda_chain_compact_dev", "dvbs2_cuda_decode_llr8", "dvbs2_cuda_decode_llr8_dev",
    "dvbs2_cuda_abi_version", "dvbs2_cuda_abi_check", "dvbs2_cuda_chain_compact_async", "dvbs2_cuda_wait",
]


class Cfg(C.Structure):
    _fields_ = [("modcod", C.c_int), ("n_ite", C.c_int), ("norm", C.c_float), ("precision", C.c_int),
                ("early_stop", C.c_int), ("max_frames", C.c_int), ("device", C.c_int), ("quant_bits", C.c_int),
                ("quant_dec", C.c_int), ("no_compaction", C.c_int), ("compaction_rounds", C.c_int),
                ("compaction_exit_eighths", C.c_int), ("compaction_nf_later", C.c_int), ("ldpc_frames_per_lane", C.c_int),
                ("ldpc_coop_max_frames", C.c_int), ("host_pipeline_chunks", C.c_int), ("ldpc_onchip", C.c_int),
                ("reserved", C.c_int * 1)]


class Info(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("K_bch", "N_bch", "K_ldpc", "N_ldpc", "bps", "n_sym", "M_ldpc", "E_ldpc",
                                       "max_frames")]


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("frames", "ldpc_iterations", "ldpc_failures", "bch_nonzero", "bch_failures",
                                          "kernel_launches")] + [("iter_hist", C.c_uint64 * 33)]


class Timers(C.Structure):
    _fields_ = ([(n + "_ms", C.c_double) for n in ("demap", "ldpc", "bch", "unpack", "pack")] +
                [(n + "_calls", C.c_uint64) for n in ("demap", "ldpc", "bch", "unpack", "pack")])


class Monitor(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("n_frames", "n_bit_errors", "n_frame_errors")]


class Dvbs2CudaError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("libdvbs2_cuda: %s (status %d)" % (msg, status))
        self.status = status


_lib = None


def load():
    """Load libdvbs2_cuda.so; raises if it has not been built (python __graft_entry__.py build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(LIB_PATH + " is missing: build it with `make -C dvbs2_b200/csrc` "
                                "(there is no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, ip = C.c_void_p, C.c_int
    L.dvbs2_cuda_default_cfg.argtypes = [C.POINTER(Cfg)]
    L.dvbs2_cuda_default_cfg.restype = None
    L.dvbs2_cuda_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    L.dvbs2_cuda_destroy.argtypes = [vp]
    L.dvbs2_cuda_get_info.argtypes = [vp, C.POINTER(Info)]
    L.dvbs2_cuda_get_stats.argtypes = [vp, C.POINTER(Stats)]
    L.dvbs2_cuda_reset_stats.argtypes = [vp]
    L.dvbs2_cuda_set_profiling.argtypes = [vp, ip]
    L.dvbs2_cuda_get_timers.argtypes = [vp, C.POINTER(Timers)]
    L.dvbs2_cuda_strerror.argtypes = [ip]
    L.dvbs2_cuda_strerror.restype = C.c_char_p
    L.dvbs2_cuda_last_error.argtypes = [vp]
    L.dvbs2_cuda_last_error.restype = C.c_char_p
    L.dvbs2_cuda_device_count.argtypes = []
    L.dvbs2_cuda_abi_version.argtypes = []
    L.dvbs2_cuda_abi_check.argtypes = [ip, C.c_size_t, C.c_size_t]
    L.dvbs2_cuda_demodulate.argtypes = [vp, vp, vp, vp, ip]
    L.dvbs2_cuda_deinterleave.argtypes = [vp, vp, vp, ip]
    L.dvbs2_cuda_ldpc_decode_siho.argtypes = [vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_bch_decode_hiho.argtypes = [vp, vp, vp, vp, ip]
    L.dvbs2_cuda_quantize.argtypes = [vp, vp, vp, C.c_size_t]
    L.dvbs2_cuda_chain.argtypes = [vp, vp, vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_host_register.argtypes = [vp, vp, C.c_size_t]
    L.dvbs2_cuda_host_unregister.argtypes = [vp, vp]
    L.dvbs2_cuda_stream.argtypes = [vp]
    L.dvbs2_cuda_stream.restype = vp
    L.dvbs2_cuda_sync.argtypes = [vp]
    L.dvbs2_cuda_demodulate_dev.argtypes = [vp, vp, vp, vp, ip, ip]
    L.dvbs2_cuda_ldpc_decode_siho_dev.argtypes = [vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_bch_decode_hiho_dev.argtypes = [vp, vp, vp, vp, ip]
    L.dvbs2_cuda_chain_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_chain_packed.argtypes = [vp, vp, vp, vp, vp, vp, vp, ip, ip]
    L.dvbs2_cuda_chain_packed_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, ip, ip]
    L.dvbs2_cuda_pl_frame_size.argtypes = [vp]
    L.dvbs2_cuda_remove_plh.argtypes = [vp, vp, vp, ip]
    L.dvbs2_cuda_remove_plh_dev.argtypes = [vp, vp, vp, ip]
    L.dvbs2_cuda_estimate.argtypes = [vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_estimate_dev.argtypes = [vp, vp, vp, vp, vp, ip]
    L.dvbs2_cuda_rx_chain.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ip, vp, ip]
    L.dvbs2_cuda_rx_chain_dev.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp, ip, vp, ip]
    L.dvbs2_cuda_iq_frame_bytes.argtypes = [vp, ip, ip]
    L.dvbs2_cuda_iq_frame_bytes.restype = C.c_size_t
    L.dvbs2_cuda_chain_compact.argtypes = [vp, ip, C.c_float, vp, vp, vp, vp, vp, vp, vp, ip, vp, ip, ip]
    L.dvbs2_cuda_chain_compact_async.argtypes = [vp, ip, C.c_float, vp, vp, vp, vp, vp, vp, vp, ip, vp, ip, ip]
    L.dvbs2_cuda_wait.argtypes = [vp]
    L.dvbs2_cuda_chain_compact_dev.argtypes = [vp, ip, C.c_float, vp, vp, vp, vp, vp, vp, vp, ip, vp, ip, ip]
    L.dvbs2_cuda_decode_llr8.argtypes = [vp, vp, vp, vp, vp, vp, ip, vp, ip]
    L.dvbs2_cuda_decode_llr8_dev.argtypes = [vp, vp, vp, vp, vp, vp, ip, vp, ip]
    L.dvbs2_cuda_tx_set_payload.argtypes = [vp, vp, ip]
    L.dvbs2_cuda_tx_random_payload.argtypes = [vp, C.c_uint64, ip]
    L.dvbs2_cuda_tx_modulate_awgn_dev.argtypes = [vp, C.c_float, C.c_uint64, vp, vp, ip]
    L.dvbs2_cuda_tx_get_bits.argtypes = [vp, ip, vp, ip]
    L.dvbs2_cuda_monitor_check.argtypes = [vp, ip]
    L.dvbs2_cuda_monitor_get.argtypes = [vp, C.POINTER(Monitor)]
    # the ctypes mirrors above must have the library's struct layout (a stale .so next to newer Python, or the reverse)
    if L.dvbs2_cuda_abi_check(ABI_VERSION, C.sizeof(Cfg), C.sizeof(Stats)) != 0:
        raise RuntimeError(LIB_PATH + " has ABI version %d, dvbs2_b200/capi.py mirrors version %d: rebuild the library"
                           % (L.dvbs2_cuda_abi_version(), ABI_VERSION))
    _lib = L
    return L


def strerror(status):
    return load().dvbs2_cuda_strerror(status).decode()
