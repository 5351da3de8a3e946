"""TEST INFRASTRUCTURE — ctypes front-ends for the CPU oracle and the compiled reference.  NOT product code.

`Oracle`  wraps oracle/libdvbs2_oracle.so (plain-C restatement, oracle/dvbs2_oracle.c).
`Ref`     wraps oracle/_ref/libdvbs2_ref.so (the unmodified reference classes, oracle/ref_harness.cpp); present only
          where oracle/build_ref.sh has run (this container; the prebuilt .so travels to the GPU box).
Only tests/, __graft_entry__.smoke() and bench.py's CPU legs import this module.
"""
import ctypes as C
import os
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
N_LDPC = 16200

# name -> bps, K_bch, N_bch(=K_ldpc), interleaver order (0 none / 1 TOP_LEFT / 2 TOP_RIGHT)  (DVBS2.cpp:287-356)
MODCODS = {
    "QPSK-S_8/9": dict(bps=2, K_bch=14232, N_bch=14400, itl=0, cstl="QPSK"),
    "QPSK-S_3/5": dict(bps=2, K_bch=9552, N_bch=9720, itl=0, cstl="QPSK"),
    "8PSK-S_3/5": dict(bps=3, K_bch=9552, N_bch=9720, itl=2, cstl="8PSK"),
    "8PSK-S_8/9": dict(bps=3, K_bch=14232, N_bch=14400, itl=1, cstl="8PSK"),
    "16APSK-S_8/9": dict(bps=4, K_bch=14232, N_bch=14400, itl=1, cstl="16APSK"),
}
_ORDER_NAME = {0: "TOP_LEFT", 1: "TOP_LEFT", 2: "TOP_RIGHT"}
# conf/mod/*.mod contents (DVB-S2 constellations, un-normalised), written to a temp file for the reference's
# Constellation_user, which only reads files.
_MOD_TEXT = {
    "QPSK": "1 1\n-1 1\n1 -1\n-1 -1\n",
    "8PSK": "0.707106781 0.707106781\n0 1\n-1 0\n-0.707106781 0.707106781\n1 0\n0.707106781 -0.707106781\n"
            "-0.707106781 -0.707106781\n0 -1\n",
    "16APSK": "0.797081141330456 0.797081141330456\n0.291751946604441 1.08883308793490\n"
              "1.08883308793490 0.291751946604441\n0.306569669742483 0.306569669742483\n"
              "-0.797081141330456 0.797081141330456\n-0.291751946604441 1.08883308793490\n"
              "-1.08883308793490 0.291751946604442\n-0.306569669742483 0.306569669742483\n"
              "0.797081141330456 -0.797081141330456\n0.291751946604441 -1.08883308793490\n"
              "1.08883308793490 -0.291751946604441\n0.306569669742483 -0.306569669742483\n"
              "-0.797081141330456 -0.797081141330456\n-0.291751946604441 -1.08883308793490\n"
              "-1.08883308793490 -0.291751946604441\n-0.306569669742483 -0.306569669742483\n",
}


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _c(a, dt):
    return np.ascontiguousarray(a, dtype=dt)


class Oracle:
    def __init__(self, path=None):
        path = path or os.path.join(HERE, "libdvbs2_oracle.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " — run `make -C oracle libdvbs2_oracle.so`")
        self.lib = L = C.CDLL(path)
        L.oracle_H_new.restype = C.c_void_p
        L.oracle_H_new.argtypes = [C.c_int]
        L.oracle_H_free.argtypes = [C.c_void_p]
        for f in ("oracle_H_M", "oracle_H_E"):
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("oracle_H_off", "oracle_H_cols"):
            getattr(L, f).argtypes = [C.c_void_p]
            getattr(L, f).restype = C.POINTER(C.c_int)
        for f in ("oracle_ldpc_f32", "oracle_ldpc_i16", "oracle_ldpc_i8"):
            getattr(L, f).argtypes = [C.c_void_p, C.c_int, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 5
        L.oracle_quantize_i16.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.oracle_quantize_i8.argtypes = [C.c_size_t, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        self._H = {}

    def H(self, K):
        if K not in self._H:
            h = self.lib.oracle_H_new(K)
            assert h, "unsupported K_ldpc %d" % K
            self._H[K] = h
        return self._H[K]

    def H_arrays(self, K):
        h = self.H(K)
        M, E = self.lib.oracle_H_M(h), self.lib.oracle_H_E(h)
        off = np.ctypeslib.as_array(self.lib.oracle_H_off(h), (M + 1,)).copy()
        cols = np.ctypeslib.as_array(self.lib.oracle_H_cols(h), (E,)).copy()
        return off, cols

    def ldpc_decode(self, K, Y, n_ite=10, norm=1.0, early_stop=True, prec="f32", want_var=False):
        dt = {"f32": np.float32, "i16": np.int16, "i8": np.int8}[prec]
        Y = _c(Y, dt).reshape(-1, N_LDPC)
        F = Y.shape[0]
        cwd = np.zeros(F, np.int8)
        V = np.zeros((F, K), np.int32)
        nit = np.zeros(F, np.int32)
        var = np.zeros((F, N_LDPC), dt) if want_var else None
        rc = getattr(self.lib, "oracle_ldpc_" + prec)(self.H(K), n_ite, norm, int(early_stop), F, _p(Y), _p(cwd), _p(V),
                                                     _p(nit), _p(var) if want_var else None)
        assert rc == 0, rc
        return (V, cwd, nit, var) if want_var else (V, cwd, nit)

    def ldpc_decode_mt(self, K, Y, threads, **kw):
        """ldpc_decode over `threads` Python threads (ctypes releases the GIL): for the >= 1e4-frame parity runs."""
        from concurrent.futures import ThreadPoolExecutor

        Y = np.asarray(Y).reshape(-1, N_LDPC)
        self.H(K)  # build the table once, before the threads race for it
        cuts = np.linspace(0, Y.shape[0], max(1, min(threads, Y.shape[0])) + 1).astype(int)
        with ThreadPoolExecutor(len(cuts) - 1) as ex:
            parts = list(ex.map(lambda i: self.ldpc_decode(K, Y[cuts[i]:cuts[i + 1]], **kw), range(len(cuts) - 1)))
        return tuple(np.concatenate([p[j] for p in parts]) for j in range(3))

    def ldpc_encode(self, K, U):
        U = _c(U, np.int32).reshape(-1, K)
        X = np.zeros((U.shape[0], N_LDPC), np.int32)
        assert self.lib.oracle_ldpc_encode(K, U.shape[0], _p(U), _p(X)) == 0
        return X

    def gf_tables(self):
        a = np.zeros(16384, np.int32)
        i = np.zeros(16384, np.int32)
        g = np.zeros(200, np.int32)
        n = C.c_int()
        self.lib.oracle_gf_tables(_p(a), _p(i), _p(g), C.byref(n))
        return a, i, g[: n.value]

    def bch_encode(self, K, N, U):
        U = _c(U, np.int32).reshape(-1, K)
        X = np.zeros((U.shape[0], N), np.int32)
        assert self.lib.oracle_bch_encode(K, N, U.shape[0], _p(U), _p(X)) == 0
        return X

    def bch_decode(self, K, N, Y):
        Y = _c(Y, np.int32).reshape(-1, N)
        F = Y.shape[0]
        cwd = np.zeros(F, np.int8)
        V = np.zeros((F, K), np.int32)
        nc = np.zeros(F, np.int32)
        assert self.lib.oracle_bch_decode(K, N, F, _p(Y), _p(cwd), _p(V), _p(nc)) == 0
        return V, cwd, nc

    def bb_scramble(self, K, U):
        U = _c(U, np.int32).reshape(-1, K)
        out = np.zeros_like(U)
        self.lib.oracle_bb_scramble(K, U.shape[0], _p(U), _p(out))
        return out

    def pl_frame_symbols(self, n_sym):
        return int(self.lib.oracle_pl_frame_symbols(int(n_sym)))

    def remove_plh(self, n_sym, Y):
        """Framer::remove_plh: [F][2 * pl_sym] -> [F][2 * n_sym]."""
        pl = self.pl_frame_symbols(n_sym)
        Y = _c(Y, np.float32).reshape(-1, 2 * pl)
        out = np.zeros((Y.shape[0], 2 * n_sym), np.float32)
        self.lib.oracle_remove_plh(int(n_sym), Y.shape[0], _p(Y), _p(out))
        return out

    def insert_plh(self, n_sym, X, seed=0):
        """Test helper (inverse of remove_plh): payload symbols into a PL frame whose header and pilot blocks hold
        arbitrary values (the receiver path never reads them)."""
        pl = self.pl_frame_symbols(n_sym)
        X = _c(X, np.float32).reshape(-1, 2 * n_sym)
        rng = np.random.default_rng(seed)
        out = rng.standard_normal((X.shape[0], 2 * pl)).astype(np.float32)
        k = np.arange(n_sym)
        pos = 90 + k + 36 * np.minimum(k // 1440, n_sym // 1440)
        out[:, 2 * pos] = X[:, 0::2]
        out[:, 2 * pos + 1] = X[:, 1::2]
        return out

    def estimate(self, n_sym, K_bch, bps, X):
        """Estimator_DVBS2::estimate: returns (sigma, Eb/N0 dB, Es/N0 dB), one value per frame."""
        X = _c(X, np.float32).reshape(-1, 2 * n_sym)
        F = X.shape[0]
        sig, eb, es = (np.zeros(F, np.float32) for _ in range(3))
        self.lib.oracle_estimate(int(n_sym), int(K_bch), int(bps), F, _p(X), _p(sig), _p(eb), _p(es))
        return sig, eb, es

    def itl_lut(self, n_cols, order):
        pi = np.zeros(N_LDPC, np.uint32)
        inv = np.zeros(N_LDPC, np.uint32)
        assert self.lib.oracle_itl_lut(n_cols, order, _p(pi), _p(inv)) == 0
        return pi, inv

    def constellation(self, bps):
        c = np.zeros(2 << bps, np.float32)
        assert self.lib.oracle_constellation(bps, _p(c)) == 0
        return c.reshape(-1, 2)

    def modulate(self, bps, X):
        X = _c(X, np.int32).reshape(-1, N_LDPC)
        sym = np.zeros((X.shape[0], 2 * N_LDPC // bps), np.float32)
        assert self.lib.oracle_modulate(bps, X.shape[0], _p(X), _p(sym)) == 0
        return sym

    def demodulate(self, bps, sigma, sym):
        sym = _c(sym, np.float32).reshape(-1, 2 * N_LDPC // bps)
        sigma = _c(sigma, np.float32).reshape(-1)
        llr = np.zeros((sym.shape[0], N_LDPC), np.float32)
        assert self.lib.oracle_demodulate(bps, sym.shape[0], _p(sigma), _p(sym), _p(llr)) == 0
        return llr

    def quantize(self, y, dec, bits, prec="i16"):
        y = _c(y, np.float32)
        out = np.zeros(y.shape, np.int16 if prec == "i16" else np.int8)
        getattr(self.lib, "oracle_quantize_" + prec)(y.size, dec, bits, _p(y), _p(out))
        return out

    # ---- whole TX chain (main.cpp:75-84 order) and AWGN, used to make synthetic frames
    def tx(self, modcod, U):
        mc = MODCODS[modcod]
        scr = self.bb_scramble(mc["K_bch"], U)
        xb = self.bch_encode(mc["K_bch"], mc["N_bch"], scr)
        xl = self.ldpc_encode(mc["N_bch"], xb)
        pi, _ = self.itl_lut(mc["bps"], mc["itl"])
        xi = xl[:, pi]
        return dict(scr=scr, bch=xb, ldpc=xl, itl=xi, sym=self.modulate(mc["bps"], xi))


def ebn0_to_sigma(modcod, ebn0_db):
    """A/src/Tools/general_utils.cpp:78-89,115-135 as used by D/mains/TX_RX_BB/main.cpp:142-148."""
    mc = MODCODS[modcod]
    rate = mc["K_bch"] / N_LDPC
    esn0 = ebn0_db + 10.0 * np.log10(rate * mc["bps"])
    return float(np.sqrt(1.0 / (2.0 * 10.0 ** (esn0 / 10.0))))


def synth_frames(oracle, modcod, n_frames, ebn0_db, seed=0xD5B52):
    """Random payload -> TX chain -> AWGN (y = c + sigma (n_re + i n_im)).  Returns dict with bits, symbols, sigma."""
    mc = MODCODS[modcod]
    rng = np.random.default_rng(seed)
    U = rng.integers(0, 2, size=(n_frames, mc["K_bch"]), dtype=np.int32)
    tx = oracle.tx(modcod, U)
    sigma = ebn0_to_sigma(modcod, ebn0_db)
    noise = rng.standard_normal(tx["sym"].shape, dtype=np.float32) * np.float32(sigma)
    tx["U"] = U
    tx["sigma"] = np.full(n_frames, sigma, np.float32)
    tx["rx"] = (tx["sym"] + noise).astype(np.float32)
    return tx


class Ref:
    """The compiled, unmodified reference.  `Ref.available()` is False on hosts where it was never built."""

    PATH = os.path.join(HERE, "_ref", "libdvbs2_ref.so")

    @classmethod
    def available(cls):
        return os.path.exists(cls.PATH)

    def __init__(self):
        self.lib = L = C.CDLL(self.PATH)
        L.ref_bch_new.restype = C.c_void_p
        L.ref_bch_new.argtypes = [C.c_int, C.c_int]
        L.ref_bch_free.argtypes = [C.c_void_p]
        L.ref_bch_decode.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        for f in ("ref_ldpc_seq_f32", "ref_ldpc_seq_i16", "ref_ldpc_seq_i8", "ref_ldpc_inter_f32", "ref_ldpc_inter_i16",
                  "ref_ldpc_tinter_f32", "ref_ldpc_tinter_i16", "ref_ldpc_seq_f32_nosyn"):
            getattr(L, f).argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.ref_ldpc_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.ref_bch_decode_mt.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 3
        L.ref_tx.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int] + [C.c_void_p] * 6
        L.ref_demodulate.argtypes = [C.c_int, C.c_char_p, C.c_char_p, C.c_int] + [C.c_void_p] * 5
        L.ref_bench_chain.restype = C.c_double
        L.ref_bench_chain.argtypes = [C.c_int, C.c_int, C.c_int, C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double,
                                      C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        self._tmp = tempfile.mkdtemp(prefix="dvbs2_mod_")
        self._mod = {}
        for k, txt in _MOD_TEXT.items():
            p = os.path.join(self._tmp, k + ".mod")
            with open(p, "w") as f:
                f.write("# " + k + "\n" + txt)
            self._mod[k] = p.encode()
        self._bch = {}

    def build_H(self, K):
        M, E = C.c_int(), C.c_int()
        self.lib.ref_build_H(K, C.byref(M), C.byref(E), None, None)
        deg = np.zeros(M.value, np.int32)
        cols = np.zeros(E.value, np.int32)
        self.lib.ref_build_H(K, C.byref(M), C.byref(E), _p(deg), _p(cols))
        return deg, cols

    def ldpc_decode(self, K, Y, n_ite=10, norm=1.0, prec="f32", simd=False, threads=1, enable_syndrome=True):
        """simd: False = sequential decoder (row a7), True = ONMS_inter (a8, what --dec-simd INTER builds),
        "tinter" = the templated ..._inter<Update_rule_NMS_simd> of C++17 builds (a9).  threads > 1 splits the frames
        over that many decoder instances."""
        dt = {"f32": np.float32, "i16": np.int16, "i8": np.int8}[prec]
        bt = {"f32": np.int32, "i16": np.int16, "i8": np.int8}[prec]
        Y = _c(Y, dt).reshape(-1, N_LDPC)
        F = Y.shape[0]
        cwd = np.zeros(F, np.int8)
        V = np.zeros((F, K), bt)
        kind = {False: "seq", True: "inter", "tinter": "tinter"}[simd]
        if not enable_syndrome:
            assert kind == "seq" and prec == "f32" and threads == 1
            rc = self.lib.ref_ldpc_seq_f32_nosyn(K, n_ite, norm, F, _p(Y), _p(cwd), _p(V))
        elif threads > 1:
            assert prec in ("f32", "i16")
            rc = self.lib.ref_ldpc_mt({"seq": 0, "inter": 1, "tinter": 2}[kind], 0 if prec == "f32" else 1, K, n_ite, norm, F,
                                      threads, _p(Y), _p(cwd), _p(V))
        else:
            rc = getattr(self.lib, "ref_ldpc_%s_%s" % (kind, prec))(K, n_ite, norm, F, _p(Y), _p(cwd), _p(V))
        assert rc >= 0, rc
        return V.astype(np.int32), cwd

    def bch_tables(self):
        a = np.zeros(16384, np.int32)
        i = np.zeros(16384, np.int32)
        g = np.zeros(200, np.int32)
        n = C.c_int()
        self.lib.ref_bch_tables(_p(a), _p(i), _p(g), C.byref(n))
        return a, i, g[: n.value]

    def bch_decode(self, K, N, Y, threads=1):
        Y = _c(Y, np.int32).reshape(-1, N)
        F = Y.shape[0]
        cwd = np.zeros(F, np.int8)
        V = np.zeros((F, K), np.int32)
        if threads > 1:
            assert self.lib.ref_bch_decode_mt(K, N, F, threads, _p(Y), _p(cwd), _p(V)) == 0
            return V, cwd
        if (K, N) not in self._bch:
            self._bch[(K, N)] = self.lib.ref_bch_new(K, N)
        self.lib.ref_bch_decode(self._bch[(K, N)], F, _p(Y), _p(cwd), _p(V))
        return V, cwd

    def close(self):
        for h in self._bch.values():
            self.lib.ref_bch_free(h)
        self._bch = {}

    def tx(self, modcod, U):
        mc = MODCODS[modcod]
        U = _c(U, np.int32).reshape(-1, mc["K_bch"])
        F = U.shape[0]
        o = dict(scr=np.zeros((F, mc["K_bch"]), np.int32), bch=np.zeros((F, mc["N_bch"]), np.int32),
                 ldpc=np.zeros((F, N_LDPC), np.int32), itl=np.zeros((F, N_LDPC), np.int32),
                 sym=np.zeros((F, 2 * N_LDPC // mc["bps"]), np.float32))
        self.lib.ref_tx(mc["K_bch"], mc["N_bch"], mc["bps"], self._mod[mc["cstl"]], _ORDER_NAME[mc["itl"]].encode(), F,
                        _p(U), _p(o["scr"]), _p(o["bch"]), _p(o["ldpc"]), _p(o["itl"]), _p(o["sym"]))
        return o

    def demodulate(self, modcod, sigma, sym):
        mc = MODCODS[modcod]
        sym = _c(sym, np.float32).reshape(-1, 2 * N_LDPC // mc["bps"])
        sigma = _c(sigma, np.float32).reshape(-1)
        F = sym.shape[0]
        itl = np.zeros((F, N_LDPC), np.float32)
        nat = np.zeros((F, N_LDPC), np.float32)
        cst = np.zeros(2 << mc["bps"], np.float32)
        self.lib.ref_demodulate(mc["bps"], self._mod[mc["cstl"]], _ORDER_NAME[mc["itl"]].encode(), F, _p(sigma), _p(sym),
                                _p(itl), _p(nat), _p(cst))
        return itl, nat, cst.reshape(-1, 2)

    def remove_plh(self, modcod, n_sym, pl_sym, Y):
        Y = _c(Y, np.float32).reshape(-1, 2 * pl_sym)
        out = np.zeros((Y.shape[0], 2 * n_sym), np.float32)
        self.lib.ref_remove_plh(modcod.encode(), int(n_sym), int(pl_sym), Y.shape[0], _p(Y), _p(out))
        return out

    def framer_generate(self, modcod, n_sym, pl_sym, X):
        X = _c(X, np.float32).reshape(-1, 2 * n_sym)
        out = np.zeros((X.shape[0], 2 * pl_sym), np.float32)
        self.lib.ref_framer_generate(modcod.encode(), int(n_sym), int(pl_sym), X.shape[0], _p(X), _p(out))
        return out

    def estimate(self, n_sym, K_bch, bps, X):
        X = _c(X, np.float32).reshape(-1, 2 * n_sym)
        F = X.shape[0]
        sig, eb, es = (np.zeros(F, np.float32) for _ in range(3))
        self.lib.ref_estimate(int(n_sym), int(K_bch), int(bps), F, _p(X), _p(sig), _p(eb), _p(es))
        return sig, eb, es

    def descramble(self, K, x):
        x = _c(x, np.int32).reshape(-1, K)
        out = np.zeros_like(x)
        self.lib.ref_descramble(K, x.shape[0], _p(x), _p(out))
        return out

    def quantize(self, y, dec, bits, prec="i16"):
        y = _c(y, np.float32)
        out = np.zeros(y.shape, np.int16 if prec == "i16" else np.int8)
        getattr(self.lib, "ref_quantize_" + prec)(int(y.size), dec, bits, _p(y), _p(out))
        return out

    def bench_chain(self, modcod, n_ite, n_threads, seconds, stages, sigma, sym, llr):
        mc = MODCODS[modcod]
        sym = _c(sym, np.float32)
        llr = _c(llr, np.float32)
        sigma = _c(sigma, np.float32)
        F = llr.reshape(-1, N_LDPC).shape[0]
        done = C.c_double()
        fps = self.lib.ref_bench_chain(mc["K_bch"], mc["N_bch"], mc["bps"], self._mod[mc["cstl"]],
                                       _ORDER_NAME[mc["itl"]].encode(), n_ite, n_threads, seconds, stages, F, _p(sigma),
                                       _p(sym), _p(llr), C.byref(done))
        return fps, done.value
