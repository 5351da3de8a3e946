"""CPU suite: the C-ABI library loads and exports what include/dvbs2_cuda.h declares; host-side logic (flags, sharding,
counter reduction over a world_size-2 gloo group); loud failure without a GPU.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _lib():
    from dvbs2_b200 import capi

    if not os.path.exists(capi.LIB_PATH):
        import subprocess

        subprocess.check_call(["make", "-j8", "-C", os.path.join(ROOT, "dvbs2_b200", "csrc")])
    return capi.load()


def test_library_exports_every_declared_symbol():
    from dvbs2_b200 import capi

    lib = _lib()
    hdr = open(os.path.join(ROOT, "include", "dvbs2_cuda.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(dvbs2_cuda_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 30
    missing = [s for s in declared if not hasattr(lib, s)]
    assert not missing, missing
    assert sorted(capi.SYMBOLS) == declared  # the ctypes binding covers the whole header


def test_structs_match_header_layout():
    from dvbs2_b200 import capi

    lib = _lib()
    assert C.sizeof(capi.Cfg) == 18 * 4 and C.sizeof(capi.Info) == 9 * 4 and C.sizeof(capi.Stats) == (6 + 33) * 8
    assert C.sizeof(capi.Timers) == 10 * 8 and C.sizeof(capi.Monitor) == 3 * 8
    cfg = capi.Cfg()
    lib.dvbs2_cuda_default_cfg(C.byref(cfg))
    assert (cfg.modcod, cfg.n_ite, cfg.norm, cfg.precision, cfg.early_stop) == (0, 10, 1.0, 0, 1)
    assert (cfg.quant_bits, cfg.quant_dec) == (6, 2)  # AFF3CT's LDPC quantiser defaults (Launcher/Code/LDPC/LDPC.cpp:58-62)
    assert lib.dvbs2_cuda_strerror(0) == b"success" and b"CUDA" in lib.dvbs2_cuda_strerror(-2)
    assert lib.dvbs2_cuda_strerror(-3) and lib.dvbs2_cuda_strerror(-99) == b"unknown status"
    # layout version: a caller built against another header is refused instead of corrupting memory
    hdr = open(os.path.join(ROOT, "include", "dvbs2_cuda.h")).read()
    assert int(re.search(r"#define DVBS2_CUDA_ABI_VERSION (\d+)", hdr).group(1)) == capi.ABI_VERSION == lib.dvbs2_cuda_abi_version()
    assert lib.dvbs2_cuda_abi_check(capi.ABI_VERSION, C.sizeof(capi.Cfg), C.sizeof(capi.Stats)) == 0
    assert lib.dvbs2_cuda_abi_check(capi.ABI_VERSION - 1, C.sizeof(capi.Cfg), C.sizeof(capi.Stats)) == -1
    assert lib.dvbs2_cuda_abi_check(capi.ABI_VERSION, C.sizeof(capi.Cfg) + 4, C.sizeof(capi.Stats)) == -1


def test_argument_validation_needs_no_gpu():
    from dvbs2_b200 import capi

    lib = _lib()
    h = C.c_void_p()
    cfg = capi.Cfg()
    lib.dvbs2_cuda_default_cfg(C.byref(cfg))
    assert lib.dvbs2_cuda_create(None, C.byref(h)) == capi.ERR_INVALID
    for field, bad in (("modcod", 7), ("n_ite", 0), ("precision", 3), ("max_frames", 0), ("norm", 1.5), ("norm", 0.0)):
        c2 = capi.Cfg.from_buffer_copy(cfg)
        setattr(c2, field, bad)
        assert lib.dvbs2_cuda_create(C.byref(c2), C.byref(h)) == capi.ERR_INVALID, field
    c2 = capi.Cfg.from_buffer_copy(cfg)
    c2.precision, c2.norm = capi.PREC["i16"], 0.3  # integer decoders only know k/8 (Update_rule_NMS.hxx:14-39)
    assert lib.dvbs2_cuda_create(C.byref(c2), C.byref(h)) == capi.ERR_INVALID
    c2 = capi.Cfg.from_buffer_copy(cfg)
    c2.ldpc_frames_per_lane = 3  # the float kernel knows 32- and 64-frame groups only
    assert lib.dvbs2_cuda_create(C.byref(c2), C.byref(h)) == capi.ERR_INVALID
    c2 = capi.Cfg.from_buffer_copy(cfg)
    c2.precision, c2.ldpc_frames_per_lane = capi.PREC["i16"], 1  # the integer precisions have one group shape
    assert lib.dvbs2_cuda_create(C.byref(c2), C.byref(h)) == capi.ERR_INVALID
    assert lib.dvbs2_cuda_sync(None) == capi.ERR_INVALID and lib.dvbs2_cuda_destroy(None) == capi.OK
    assert lib.dvbs2_cuda_pl_frame_size(None) == capi.ERR_INVALID


def test_no_cpu_fallback_without_a_device():
    """On a host without a CUDA device the product path must fail loudly, never fall back to the oracle."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("this host has a GPU")
    from dvbs2_b200 import Handle, capi

    _lib()
    assert capi.load().dvbs2_cuda_device_count() == 0
    with pytest.raises(capi.Dvbs2CudaError) as e:
        Handle()
    assert e.value.status == capi.ERR_CUDA
    import dvbs2_b200

    src = "".join(open(os.path.join(ROOT, "dvbs2_b200", f)).read() for f in os.listdir(os.path.join(ROOT, "dvbs2_b200"))
                  if f.endswith(".py"))
    assert "import oracle" not in src and "from oracle" not in src and dvbs2_b200.__name__ == "dvbs2_b200"


def test_factory_flags_follow_the_reference():
    from dvbs2_b200.factory import DVBS2

    p = DVBS2(["--mod-cod", "8PSK-S_3/5", "--dec-implem", "NMS", "--dec-ite", "10", "--dec-simd", "INTER", "-F", "16"])
    assert (p.modcod, p.ldpc_implem, p.ldpc_nite, p.n_frames, p.norm, p.precision) == ("8PSK-S_3/5", "NMS", 10, 16, 1.0, "f32")
    assert DVBS2([]).ldpc_nite == 50  # DVBS2.cpp:135 default
    assert DVBS2(["--dec-implem", "MS", "--dec-norm", "0.75"]).norm == 1.0
    assert DVBS2(["--dec-implem", "NMS", "--dec-norm", "0.75", "--dec-prec", "16"]).precision == "i16"
    with pytest.raises(ValueError):
        DVBS2(["--dec-implem", "SPA"])
    with pytest.raises(SystemExit):
        DVBS2(["--mod-cod", "32APSK-S_9/10"])
    with pytest.raises(ValueError):
        DVBS2(["-F", "0"])


def test_sigma_formula_matches_oracle():
    from dvbs2_b200 import synth
    from oracle.oracle import MODCODS, ebn0_to_sigma

    for m in MODCODS:
        for e in (0.0, 2.2, 4.0, 7.5):
            assert abs(synth.ebn0_to_sigma(m, e) - ebn0_to_sigma(m, e)) < 1e-12
        assert synth.MODCODS[m]["K_bch"] == MODCODS[m]["K_bch"] and synth.MODCODS[m]["bps"] == MODCODS[m]["bps"]
    assert abs(synth.ebn0_to_sigma("QPSK-S_8/9", 4.0) - 0.3366) < 1e-4


def test_shard_frames():
    from dvbs2_b200.dist import shard_frames

    for n in (0, 1, 7, 8, 100, 75776):
        for w in (1, 2, 3, 8):
            parts = [shard_frames(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and sum(c for _, c in parts) == n
            assert all(parts[i][0] + parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in parts) - min(c for _, c in parts) <= 1
    with pytest.raises(ValueError):
        shard_frames(10, 2, 2)


def _rank_main(rank, world, port, out):
    import torch.distributed as dist

    from dvbs2_b200.dist import Monitor_BFER_reduction, reduce_counters, reduce_max, shard_frames

    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    first, cnt = shard_frames(1001, rank, world)
    # each rank "decodes" its shard: frame f has f % 7 bit errors
    f = np.arange(first, first + cnt)
    local = {"n_frames": cnt, "n_be": int((f % 7).sum()), "n_fe": int((f % 7 != 0).sum())}
    tot = reduce_counters(local)
    mon = Monitor_BFER_reduction(max_fe=800)
    mon.update(local)
    t = reduce_max(0.5 + rank)
    # the Monte-Carlo driver's stop rule: every rank runs batches until the GLOBAL frame-error count reaches max_fe
    from dvbs2_b200.montecarlo import run_point

    cum = {"n_frames": 0, "n_be": 0, "n_fe": 0}

    def step(b):
        cum["n_frames"] += 100
        cum["n_be"] += 10
        cum["n_fe"] += 3 if rank == 0 else 1  # 4 frame errors per batch over both ranks
        return dict(cum)

    red, n_batches = run_point(step, max_fe=20)
    assert n_batches == 5 and red.total == {"n_frames": 1000, "n_be": 100, "n_fe": 20}, (n_batches, red.total)
    red, n_batches = run_point(step, max_fe=10 ** 9, max_frames=2500)
    assert red.total["n_frames"] >= 2500 and red.is_done()
    out.put((rank, tot, mon.is_done(), mon.rates(10), t))
    dist.destroy_process_group()


def test_counter_reduction_world_size_2_gloo():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + os.getpid() % 2000
    ps = [ctx.Process(target=_rank_main, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(60)
        assert p.exitcode == 0
    f = np.arange(1001)
    want = {"n_be": int((f % 7).sum()), "n_fe": int((f % 7 != 0).sum()), "n_frames": 1001}
    for rank, tot, done, rates, t in res:
        assert tot == want and done and t == 1.5
        assert abs(rates["fer"] - want["n_fe"] / 1001) < 1e-12 and abs(rates["ber"] - want["n_be"] / 10010) < 1e-12


def test_committed_bench_line_follows_the_driver_contract():
    """The bench line recorded from the final code of the round (profiles/r2_final_bench.json) carries every key the driver and
    the judge read, with consistent values."""
    import json

    d = json.load(open(os.path.join(ROOT, "profiles", "r2_final_bench.json")))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "Mbit/s" and d["higher_is_better"] is True and d["scaling"] == "weak" and d["vs_baseline"] is None
    assert d["dtype"] == "f32" and "workload" in d["config"] and d["gpu_launches"] > 0
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and abs(r["frac"] - r["achieved"] / r["peak"]) < 1e-4 and r["traffic"] > 0
    c = d["cpu_baseline"]
    assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and c["sample"]
    e = d["e2e"]
    assert e["value"] > 0 and e["h2d_bytes_per_step"] > 0 and e["d2h_bytes_per_step"] > 0 and e["matches_device_path"] is True
    assert e["value"] < d["value"]  # host buffers over PCIe cannot beat the device-resident rate
    assert set(e["sockets"]) >= {"f32_int32", "sc16_packed"} and e["headline_sockets"] == "sc16_packed"
    for name, line in d["configs"].items():  # BASELINE configs 2-4 are lines of the same object
        assert line["value"] > 0 and line["roofline"]["frac"] > 0, name
    assert d["int16"]["value"] > d["value"] and sum(d["iter_hist"]) == d["fer"]["frames"]
    K_bch = 14232
    frames = d["config"]["frames_per_step_per_gpu"] * d["steps"] * d["n_gpus"]
    assert abs(d["value"] - K_bch * frames / (d["ms_per_step"] * d["steps"] * 1e-3) / 1e6) / d["value"] < 1e-3
    ref = json.load(open(os.path.join(ROOT, "profiles", "r2_final_bench_reference_arm.json")))
    assert ref["impl"] == "reference" and ref["metric"] == d["metric"] and ref["unit"] == d["unit"]
    assert ref["e2e"]["h2d_bytes_per_step"] == 0 and ref["cpu_baseline"]["kind"] == "reference"


def test_sc12_wire_format_round_trip():
    """12-bit samples, three bytes per symbol (DVBS2_IQ_SC12): I = bits 0..11, Q = bits 12..23 of a little-endian word —
    pack / unpack of the Python mirror are inverse up to half an LSB, saturate at +-2047 and keep the sign."""
    from dvbs2_b200.modules import iq_pack, iq_unpack

    rng = np.random.default_rng(3)
    x = (rng.standard_normal((7, 32)) * 1.3).astype(np.float32)
    x[0, :4] = [5.0, -5.0, 0.0, -1.0 / 512]  # beyond full scale (+-4.0 at LSB 2^-9), zero, one negative LSB
    w = iq_pack(x, "sc12", 1.0 / 512)
    assert w.dtype == np.uint8 and w.shape == (7, 48)
    y = iq_unpack(w, "sc12", 1.0 / 512)
    assert y.shape == x.shape
    inside = np.abs(x) < 3.99
    assert np.abs(y - x)[inside].max() <= 0.5 / 512 + 1e-7
    assert y[0, 0] == 2047 / 512 and y[0, 1] == -2047 / 512 and y[0, 2] == 0.0 and y[0, 3] == -1.0 / 512
    # byte layout of one symbol: I = 0x123, Q = -2 (0xffe)
    one = iq_pack(np.array([[0x123 / 512, -2 / 512]], np.float32), "sc12", 1.0 / 512)
    assert one.tolist() == [[0x23, 0xe1, 0xff]]
