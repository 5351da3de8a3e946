"""Flag surface of factory::DVBS2 for the decode path (src/common/Factory/DVBS2/DVBS2.cpp:60-203, 287-356).

Parses the same command-line options the reference binaries take for this path and builds the GPU modules:
  --mod-cod {QPSK-S_8/9,QPSK-S_3/5,8PSK-S_3/5,8PSK-S_8/9,16APSK-S_8/9}   (DVBS2.cpp:67-69)
  --dec-implem {MS,NMS}   (DVBS2.cpp:79-81; SPA is a CPU-only implementation and is rejected)
  --dec-ite N             (DVBS2.cpp:76-78, default 50)
  --dec-simd {INTER,INTRA}  accepted and ignored: the device path is always inter-frame parallel
  -F / --sim-inter-fra N  (DVBS2.cpp:70-72)  frames per task call
plus the GPU-only knobs SURVEY.md §5 proposes: --dec-prec {32,16,8}, --dec-norm, --gpu-id, --dec-no-synd.
As in the reference --dec-implem NMS runs with normalisation 1.0 unless --dec-norm is given.
"""
import argparse

from .modules import (Decoder_BCH_DVBS2_cuda, Decoder_chain_DVBS2_cuda, Decoder_LDPC_DVBS2_cuda, Handle,
                      Interleaver_DVBS2_cuda, Modem_DVBS2_cuda)

MODCODS = ["QPSK-S_8/9", "QPSK-S_3/5", "8PSK-S_3/5", "8PSK-S_8/9", "16APSK-S_8/9"]


class DVBS2:
    def __init__(self, argv=None):
        p = argparse.ArgumentParser(prog="dvbs2_b200", add_help=True)
        p.add_argument("--mod-cod", default="QPSK-S_8/9", choices=MODCODS)
        p.add_argument("--dec-implem", default="NMS", choices=["MS", "NMS", "SPA"])
        p.add_argument("--dec-ite", type=int, default=50)
        p.add_argument("--dec-simd", default="", choices=["", "INTER", "INTRA"])
        p.add_argument("-F", "--sim-inter-fra", dest="n_frames", type=int, default=1)
        p.add_argument("--dec-prec", type=int, default=32, choices=[32, 16, 8])
        p.add_argument("--dec-norm", type=float, default=1.0)
        p.add_argument("--dec-no-synd", action="store_true")
        p.add_argument("--gpu-id", type=int, default=0)
        a = p.parse_args(argv)
        if a.dec_implem == "SPA":
            raise ValueError("--dec-implem SPA is not available on the GPU path (MS / NMS only)")
        if a.dec_ite <= 0 or a.n_frames <= 0:
            raise ValueError("--dec-ite and -F have to be strictly positive")
        self.modcod = a.mod_cod
        self.ldpc_implem = a.dec_implem
        self.ldpc_nite = a.dec_ite
        self.ldpc_simd = a.dec_simd
        self.n_frames = a.n_frames
        self.precision = {32: "f32", 16: "i16", 8: "i8"}[a.dec_prec]
        self.norm = 1.0 if a.dec_implem == "MS" else a.dec_norm
        self.early_stop = not a.dec_no_synd
        self.device = a.gpu_id

    def build_handle(self):
        return Handle(modcod=self.modcod, n_ite=self.ldpc_nite, norm=self.norm, precision=self.precision,
                      early_stop=self.early_stop, n_frames=self.n_frames, device=self.device)

    # same names as factory::DVBS2::build_* (DVBS2.cpp:395-488)
    def build_modem(self, handle):
        return Modem_DVBS2_cuda(handle, self.n_frames)

    def build_itl(self, handle):
        return Interleaver_DVBS2_cuda(handle, self.n_frames)

    def build_ldpc_decoder(self, handle):
        return Decoder_LDPC_DVBS2_cuda(handle, self.n_frames)

    def build_bch_decoder(self, handle):
        return Decoder_BCH_DVBS2_cuda(handle, self.n_frames)

    def build_chain(self, handle):
        return Decoder_chain_DVBS2_cuda(handle, self.n_frames)
