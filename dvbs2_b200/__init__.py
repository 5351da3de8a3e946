"""dvbs2_b200 — B200-native DVB-S2 baseband decode chain (demapper, LDPC NMS, BCH) behind libdvbs2_cuda.

Only the hot path lives here: csrc/ (CUDA kernels + the C ABI) and the host-side mirror of the reference's module
interface.  Nothing in this package imports oracle/.
"""
from . import capi  # noqa: F401
from .modules import (Decoder_BCH_DVBS2_cuda, Decoder_chain_compact_DVBS2_cuda, Decoder_chain_DVBS2_cuda,  # noqa: F401
                      Decoder_LDPC_DVBS2_cuda, Decoder_llr8_DVBS2_cuda,
                      Estimator_DVBS2_cuda, Framer_DVBS2_cuda, Handle, Interleaver_DVBS2_cuda, Modem_DVBS2_cuda,
                      Receiver_chain_DVBS2_cuda, Transmitter_DVBS2_cuda)

__all__ = ["capi", "Handle", "Modem_DVBS2_cuda", "Interleaver_DVBS2_cuda", "Decoder_LDPC_DVBS2_cuda",
           "Decoder_BCH_DVBS2_cuda", "Decoder_chain_DVBS2_cuda", "Transmitter_DVBS2_cuda", "Framer_DVBS2_cuda",
           "Estimator_DVBS2_cuda", "Receiver_chain_DVBS2_cuda", "Decoder_chain_compact_DVBS2_cuda", "Decoder_llr8_DVBS2_cuda"]
