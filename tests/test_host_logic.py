"""CPU-side checks of the product's own host-testable pieces:
  * the device field / curve headers (fp.cuh, ec.cuh) compiled for the host with the bit-exact
    emulation of the PTX carry chain, compared with the oracle;
  * the index / control-logic models of the NTT and MSM kernels (tests/models);
  * the C ABI: the library loads, exports every symbol include/summa_b200.h declares, and refuses
    to create a context without a GPU (no CPU fallback)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bn254 as B
from tests.models import msm_model, ntt_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("ht") / "ht.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "host_field_harness.cpp")])
    return ctypes.CDLL(so)


def _arr(bs):
    return np.frombuffer(bs, dtype=np.uint32).copy()


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


@pytest.mark.parametrize("name,mod", [("fr", B.R), ("fq", B.Q)])
def test_device_field_code_on_host(harness, name, mod):
    fn = getattr(harness, f"ht_{name}_op")
    rnd = random.Random(7)
    edge = [0, 1, mod - 1, mod - 2, (1 << 256) % mod, 2, (1 << 254) % mod, mod >> 1]
    xs = edge + [rnd.randrange(mod) for _ in range(2000)]
    ys = list(reversed(edge)) + [rnd.randrange(mod) for _ in range(2000)]
    a = _arr(b"".join(x.to_bytes(32, "little") for x in xs))
    b = _arr(b"".join(y.to_bytes(32, "little") for y in ys))
    out = np.empty_like(a)
    n = len(xs)
    rinv = pow(1 << 256, -1, mod)
    dec = lambda m: [int.from_bytes(out[8 * i:8 * i + 8].tobytes(), "little") for i in range(m)]
    fn(_p(out), _p(a), _p(b), ctypes.c_size_t(n), 0)
    assert dec(n) == [x * y * rinv % mod for x, y in zip(xs, ys)]
    fn(_p(out), _p(a), _p(b), ctypes.c_size_t(n), 1)
    assert dec(n) == [(x + y) % mod for x, y in zip(xs, ys)]
    fn(_p(out), _p(a), _p(b), ctypes.c_size_t(n), 2)
    assert dec(n) == [(x - y) % mod for x, y in zip(xs, ys)]
    m = 20
    fn(_p(out), _p(a[8 * 8:]), _p(b), ctypes.c_size_t(m), 3)
    assert dec(m) == [pow(x, -1, mod) * (1 << 512) % mod for x in xs[8:8 + m]]


def test_device_curve_code_on_host(harness):
    rnd = random.Random(9)
    pts = [B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R)) for _ in range(40)]
    pts[3] = None
    pts[10] = pts[11]
    pts[20] = B.g1_neg(pts[21])
    pts[30] = pts[2]
    negs = np.array([rnd.randrange(2) for _ in pts], dtype=np.uint8)
    negs[10] = negs[11] = negs[20] = negs[21] = 0
    ref = None
    for p, ng in zip(pts, negs):
        ref = B.g1_add(ref, B.g1_neg(p) if ng else p)
    out = np.zeros(16, dtype=np.uint32)
    harness.ht_sum_points(_p(out), _p(_arr(B.g1s_to_bytes(pts))), _p(negs), ctypes.c_size_t(len(pts)))
    assert B.g1_from_mont_bytes(out.tobytes()) == ref
    for case in ([pts[0], pts[0]], [pts[0], B.g1_neg(pts[0])], [None, pts[0]], [pts[0], None], [None, None]):
        harness.ht_sum_points(_p(out), _p(_arr(B.g1s_to_bytes(case + case))), _p(np.zeros(4, dtype=np.uint8)), ctypes.c_size_t(4))
        ref = None
        for p in case + case:
            ref = B.g1_add(ref, p)
        assert B.g1_from_mont_bytes(out.tobytes()) == ref
    harness.ht_double_xyzz(_p(out), _p(_arr(B.g1_to_mont_bytes(pts[1]))), 5)
    assert B.g1_from_mont_bytes(out.tobytes()) == B.g1_mul(pts[1], 32)
    harness.ht_add_xyzz(_p(out), _p(_arr(B.g1_to_mont_bytes(pts[1]))), _p(_arr(B.g1_to_mont_bytes(B.g1_mul(pts[1], 2)))))
    assert B.g1_from_mont_bytes(out.tobytes()) == B.g1_mul(pts[1], 8)  # 4p + 2(2p): general-add doubling path


@pytest.mark.parametrize("tile_log,rmax_log,ks", [(4, 3, range(1, 12)), (11, 8, (11, 12, 13))])
def test_ntt_plan_model_matches_best_fft(tile_log, rmax_log, ks):
    rnd = random.Random(3)
    old = ntt_model.TILE_LOG, ntt_model.RMAX_LOG
    ntt_model.TILE_LOG, ntt_model.RMAX_LOG = tile_log, rmax_log
    try:
        for k in ks:
            w = B.omega_for(k)
            a = [rnd.randrange(B.R) for _ in range(1 << k)]
            assert ntt_model.ntt(a, w, k) == B.best_fft(a, w, k), (k, ntt_model.plan(k))
    finally:
        ntt_model.TILE_LOG, ntt_model.RMAX_LOG = old


def _scalars(rnd, n, mode):
    sc = [rnd.randrange(B.R) for _ in range(n)]
    if mode == "Z":
        sc = [s if rnd.random() < 0.1 else 0 for s in sc]
    if mode == "C":
        v = rnd.randrange(B.R)
        sc = [v if rnd.random() < 0.9 else s for s in sc]
    if mode == "S":
        sc = [rnd.randrange(256) for _ in sc]
    return sc


def test_msm_pipeline_model():
    rnd = random.Random(5)
    g = msm_model.IntGroup(B.R)
    for trial in range(120):
        n = rnd.choice([1, 2, 3, 7, 16, 33, 100, 257])
        c = rnd.choice([2, 3, 4, 5, 8])
        mode = rnd.choice("UZCS")
        sc = _scalars(rnd, n, mode)
        if trial % 17 == 0:
            sc[0] = B.R - 1
        bs = [rnd.randrange(B.R) for _ in range(n)]
        exp = sum(s * b for s, b in zip(sc, bs)) % B.R
        got = msm_model.msm(g, sc, bs, c, L1=rnd.choice([4, 8]), LK=rnd.choice([3, 4]), final_max=rnd.choice([4, 16]), seg_log=rnd.choice([0, 1, 2, 3]))
        assert got == exp, (trial, n, c, mode)


def test_msm_model_tables_batches_and_window_shards():
    """fixed-base window tables (one shared bucket set), batches of scalar vectors in one launch set, and the window-range shards of
    the multi-GPU proof: partial sums over disjoint window ranges add up to the commitment (with tables) or fold by Horner (without)."""
    rnd = random.Random(11)
    g = msm_model.IntGroup(B.R)
    for trial in range(40):
        n = rnd.choice([1, 5, 16, 40, 129])
        c = rnd.choice([3, 4, 6, 9])
        batch = rnd.choice([1, 2, 3, 5])
        bs = [rnd.randrange(B.R) for _ in range(n)]
        sc = []
        for _ in range(batch):
            sc += _scalars(rnd, n, rnd.choice("UZCS"))
        exp = [sum(s * b for s, b in zip(sc[j * n:(j + 1) * n], bs)) % B.R for j in range(batch)]
        kw = dict(L1=rnd.choice([4, 8]), LK=4, final_max=rnd.choice([4, 16]), seg_log=rnd.choice([1, 2, 3]), cta_scan_max=rnd.choice([0, 64, 1 << 20]))
        assert msm_model.msm(g, sc, bs, c, tables=True, batch=batch, **kw) == exp, (trial, "batch")
        W = msm_model.window_count(c)
        world = rnd.choice([2, 4, 8])
        if W >= world:
            parts = [msm_model.msm(g, sc, bs, c, tables=True, batch=batch, w_lo=r * W // world, w_hi=(r + 1) * W // world, **kw) for r in range(world)]
            assert [sum(p[j] for p in parts) % B.R for j in range(batch)] == exp, (trial, "window shards, tables")
            wins = []
            for r in range(world):
                wins += msm_model.msm(g, sc[:n], bs, c, w_lo=r * W // world, w_hi=(r + 1) * W // world, **kw)
            acc = 0
            for w in reversed(range(W)):
                acc = (acc * (1 << c) + wins[w]) % B.R
            assert acc == exp[0], (trial, "window shards, plain bases")


def test_abi_library_loads_and_exports_every_declared_symbol():
    from circuits_halo2_b200 import _lib
    L = _lib.lib()
    names = _lib.declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/summa_b200.h but not exported"
    assert L.sb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device, context creation must fail loudly (SB_ERR_NO_DEVICE)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible; the refusal path is exercised on the CPU box")
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200._lib import SummaB200Error
    with pytest.raises(SummaB200Error) as e:
        sb.Context(0)
    assert e.value.status == 3


@pytest.mark.parametrize("seed", [1, 2, 3, 99])
def test_evaluate_h_compiler_matches_direct_evaluation(seed):
    """The evaluate_h compiler (global CSE across the 29 terms, accumulation order chosen for short live ranges, product-free
    sub-expressions recomputed per term) must compute exactly sum_i y^(T-1-i) term_i: the host interpreter of the compiled
    program and a direct walk of the expression trees agree on a pseudo-random row; the program keeps the shape the kernel's
    occupancy was tuned for."""
    import ctypes
    from circuits_halo2_b200 import _lib
    L = _lib.lib()
    cs = open(os.path.join(ROOT, "tests", "golden", "mst_inclusion_cs.json")).read().encode()
    a, b = np.zeros(4, dtype=np.uint64), np.zeros(4, dtype=np.uint64)
    shape = (ctypes.c_uint32 * 4)()
    _lib.check(L.sb_test_h_program(ctypes.c_char_p(cs), ctypes.c_uint64(seed), a.ctypes.data_as(ctypes.c_void_p), b.ctypes.data_as(ctypes.c_void_p), shape), "sb_test_h_program")
    assert a.any() and (a == b).all()
    instr, n_mul, n_add, slots = list(shape)
    assert n_mul <= 125 and slots <= 8 and n_mul + n_add <= instr <= n_mul + n_add + 1, list(shape)
