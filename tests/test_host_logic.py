"""CPU-side checks of the product's own host-testable pieces:
  * the device field / curve headers (fp.cuh, ec.cuh) compiled for the host with the bit-exact
    emulation of the PTX carry chain, compared with the oracle;
  * the NTT kernel's tile code (csrc/ntt_core.cuh) run for every thread of every tile on the host, and the control-logic model of the
    MSM kernels (tests/models);
  * the C ABI: the library loads, exports every symbol include/summa_b200.h declares, and refuses
    to create a context without a GPU (no CPU fallback)."""
import ctypes
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bn254 as B
from tests.models import msm_model

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
    # the dedicated squaring, with limb patterns that saturate its carry chains
    pat = [0xFFFFFFFF, 0, 1, 0x80000000, 0xFFFFFFFE, 0x7FFFFFFF]
    sq = list(xs)
    for t in range(3000):
        v = 0
        for limb in range(8):
            v |= (rnd.choice(pat) if rnd.random() < 0.7 else rnd.getrandbits(32)) << (32 * limb)
        sq.append(v % mod)
    a2 = _arr(b"".join(x.to_bytes(32, "little") for x in sq))
    out2 = np.empty_like(a2)
    fn(_p(out2), _p(a2), _p(a2), ctypes.c_size_t(len(sq)), 7)
    got = [int.from_bytes(out2[8 * i:8 * i + 8].tobytes(), "little") for i in range(len(sq))]
    assert got == [x * x * rinv % mod for x in sq]
    m = 20
    fn(_p(out), _p(a[8 * 8:]), _p(b), ctypes.c_size_t(m), 6)   # Fermat
    assert dec(m) == [pow(x, -1, mod) * (1 << 512) % mod for x in xs[8:8 + m]]
    m = 1000                                                    # binary extended Euclid (the one in use), edge values included; inv(0) = 0
    fn(_p(out), _p(a), _p(b), ctypes.c_size_t(m), 3)
    assert dec(m) == [pow(x, -1, mod) * (1 << 512) % mod if x else 0 for x in xs[:m]]


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
        # tree_log: buckets per CTA of the tree bucket reduction (2^8 in CUDA); 1 and 2 reach the two-stage and the running-sum-first paths at these sizes
        got = msm_model.msm(g, sc, bs, c, L1=rnd.choice([4, 8]), LK=rnd.choice([3, 4]), final_max=rnd.choice([4, 16]), seg_log=rnd.choice([0, 1, 2, 3]),
                            tree_log=rnd.choice([None, 1, 2, 3, 8]))
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
        kw = dict(L1=rnd.choice([4, 8]), LK=4, final_max=rnd.choice([4, 16]), seg_log=rnd.choice([1, 2, 3]), cta_scan_max=rnd.choice([0, 64, 1 << 20]),
                  tree_log=rnd.choice([None, 1, 2, 3, 8]))
        assert msm_model.msm(g, sc, bs, c, tables=True, batch=batch, **kw) == exp, (trial, "batch")
        W = msm_model.window_count(c)
        world = rnd.choice([2, 4, 8])
        lg = world.bit_length() - 1
        if c >= lg + 2:   # bucket-residue shards (the multi-GPU split of table MSMs): partial sums over the residues add up
            kw_t = dict(kw, tree_log=kw["tree_log"] or 2)
            parts = [msm_model.msm(g, sc, bs, c, tables=True, batch=batch, residue=(r, lg), **kw_t) for r in range(world)]
            assert [sum(p[j] for p in parts) % B.R for j in range(batch)] == exp, (trial, "residue shards")
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


# ------------------------------------------------------------------ the NTT tile code (csrc/ntt_core.cuh) on the host
@pytest.fixture(scope="module")
def ntt_harness(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("htntt") / "htntt.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", so,
                           os.path.join(ROOT, "tests", "host", "host_ntt_harness.cpp")])
    lib = ctypes.CDLL(so)
    lib.ht_ntt.restype = ctypes.c_uint32
    return lib


def _run_device_ntt_on_host(lib, log_n, tile, min_passes, full_tw, seed=0, scale=None, pre_vec=None, pre_pat=None, post_pat=None, post_vec=None, n_in=0, n_out=0, eb=3):
    """Every thread of every tile of every pass of ntt_pass_kernel, emulated phase by phase on the CPU, vs the oracle's best_fft
    with the same fused scalings applied as separate passes.  Returns (equal, worst shared-memory bank-conflict degree)."""
    from oracle import cpu
    n = 1 << log_n
    a = cpu.random_fr(n, seed + log_n)
    w = np.frombuffer(B.fr_to_mont_bytes(B.omega_for(log_n)), dtype=np.uint64).copy()
    got = a.copy()
    pp = lambda x: x.ctypes.data_as(ctypes.c_void_p) if x is not None else None
    worst = lib.ht_ntt(pp(got), pp(w), log_n, tile, min_passes, int(full_tw), pp(scale), pp(pre_vec), pp(pre_pat), 0 if pre_pat is None else pre_pat.shape[0],
                       pp(post_pat), 0 if post_pat is None else post_pat.shape[0], ctypes.c_uint64(n_in), ctypes.c_uint64(n_out), pp(post_vec), eb)
    x = a.copy()
    if n_in:
        x[n_in:] = 0
    if pre_vec is not None:
        x = cpu.fr_mul(x.reshape(-1), pre_vec.reshape(-1)).reshape(-1, 4)
    if pre_pat is not None:
        x = cpu.fr_scale_pattern(x.reshape(-1), pre_pat.reshape(-1)).reshape(-1, 4)
    ref = cpu.best_fft(x.reshape(-1), w, log_n, threads=4).reshape(-1, 4)
    if scale is not None:
        ref = cpu.fr_scale(ref.reshape(-1), scale).reshape(-1, 4)
    if post_pat is not None:
        ref = cpu.fr_scale_pattern(ref.reshape(-1), post_pat.reshape(-1)).reshape(-1, 4)
    if post_vec is not None:
        ref = cpu.fr_mul(ref.reshape(-1), post_vec.reshape(-1)).reshape(-1, 4)
    no = n_out or n
    return bool((got[:no] == ref[:no]).all() and (got[no:].view(np.uint8) == 0xAB).all()), worst


@pytest.mark.parametrize("tile", [10, 11, 12])
def test_device_ntt_tile_code_on_host_all_plans(ntt_harness, tile):
    """single pass (2^tile), two passes and three passes, full inter-pass twiddle tables and the two-level fallback; every shared-memory
    access pattern (exchanges, twiddle reads, the last pass' transposition) at most 2-way bank conflicted."""
    for log_n in range(tile, 17):
        for min_passes in (0, 3):
            for full_tw in (0, 1):
                if log_n <= tile and (min_passes or full_tw):
                    continue
                for eb in (3, 2):   # eight (radix-8 stages) or four (radix-4 stages) elements per thread
                    ok, worst = _run_device_ntt_on_host(ntt_harness, log_n, tile, min_passes, full_tw, eb=eb)
                    assert ok, (tile, log_n, min_passes, full_tw, eb)
                    assert worst <= 2, (tile, log_n, min_passes, full_tw, eb, worst)


@pytest.mark.parametrize("log_n,tile,min_passes,full_tw", [(11, 11, 0, 0), (12, 12, 0, 0), (13, 11, 0, 1), (15, 11, 3, 0), (15, 12, 0, 1)])
def test_device_ntt_fused_scalings_on_host(ntt_harness, log_n, tile, min_passes, full_tw):
    """what EvaluationDomain folds into the transform: n^-1 (scale), zeta^(i mod 3) on the way in, t(X)^-1 (8-pattern) on the way in, zeta^-(i mod 3) on
    the way out, vector scalings (coset powers, g^-i / n), zero padding (n_in) and truncation (n_out) -- one at a time and all together"""
    from oracle import cpu
    n = 1 << log_n
    m = lambda x: np.frombuffer(B.fr_to_mont_bytes(x % B.R), dtype=np.uint64).copy()
    scale, pre_vec, post_vec = m(pow(n, -1, B.R)), cpu.random_fr(n, 77), cpu.random_fr(n, 80)
    pat8, pat3 = cpu.random_fr(8, 78), cpu.random_fr(3, 79)
    cases = [dict(scale=scale), dict(pre_vec=pre_vec), dict(pre_pat=pat8), dict(pre_pat=pat3), dict(post_pat=pat3), dict(post_vec=post_vec), dict(n_in=n // 8),
             dict(n_out=5 * n // 8), dict(scale=scale, pre_vec=pre_vec, pre_pat=pat3, post_pat=pat3, post_vec=post_vec, n_in=n // 2, n_out=n // 2 + 3)]
    for kw in cases:
        ok, _ = _run_device_ntt_on_host(ntt_harness, log_n, tile, min_passes, full_tw, **kw)
        assert ok, (log_n, tile, sorted(kw))


# ------------------------------------------------------------------ witness generation (csrc/witness.cpp, host C++ of the product)
def _product_witness(proof, levels, nc, k, n_bytes=8):
    import circuits_halo2_b200 as sb
    pre = [proof["entry"].preimage(), proof["sibling_leaf_node_hash_preimage"]] + list(proof["sibling_middle_node_hash_preimages"])
    flat = np.concatenate([np.frombuffer(B.fr_to_mont_bytes(v % B.R), dtype=np.uint64) for p in pre for v in p])
    inst, cells, vals = sb.mst_inclusion_witness(levels, nc, k, flat, np.array(proof["path_indices"], dtype=np.uint8), n_bytes)
    return inst, {(int(c), int(r)): B.fr_from_mont_bytes(v.tobytes()) for (c, r), v in zip(cells, vals)}


def test_product_witness_equals_oracle_synthesis():
    """sb_mst_inclusion_witness (the product's restatement of `MstInclusionCircuit::synthesize`'s witness side + SimpleFloorPlanner placement) vs
    the oracle's synthesis (oracle/mst_circuit.py, pinned on the reference vk and MockProver region indices): every advice cell and every instance,
    for LEVELS = 4 (csv/entry_16.csv, several users), 3 currencies with ragged padding, and the LEVELS = 20 / LEVELS = 23 x 8-currency fixtures."""
    from oracle import mst as M
    from oracle import mst_circuit as C
    golden = os.path.join(ROOT, "tests", "golden")
    tree = M.MerkleSumTree.from_csv(os.path.join(golden, "entry_16.csv"))
    for user in (0, 5, 15):
        pr = tree.generate_proof(user)
        lay = C.synthesize(11, pr, 4, 2, 8)
        want = {(col, row): v for col, dense in enumerate(C.advice_columns(lay)) for row, v in enumerate(dense) if v}
        inst, got = _product_witness(pr, 4, 2, 11)
        assert got == want and inst == [tree.nodes[0][user][0], tree.root[0]] + tree.root[1]
    t3 = M.MerkleSumTree([M.Entry(f"u{i}", [i * 1000 + 7, 5 * i, (1 << 60) + i]) for i in range(5)])
    pr = t3.generate_proof(3)
    lay = C.synthesize(12, pr, 3, 3, 8)
    inst, got = _product_witness(pr, 3, 3, 12)
    assert got == {(col, row): v for col, dense in enumerate(C.advice_columns(lay)) for row, v in enumerate(dense) if v}
    assert inst == [t3.nodes[0][3][0], t3.root[0]] + t3.root[1]
    um = lambda x: B.fr_from_mont_bytes(np.ascontiguousarray(x).tobytes())

    class _E:
        def __init__(self, p):
            self.p = p

        def preimage(self):
            return self.p
    for name, levels, nc, k in (("mst_inclusion_assignment_l20_tree.npz", 20, 2, 13), ("mst_inclusion_assignment_l23_n8_tree.npz", 23, 8, 15)):
        fx = np.load(os.path.join(golden, name))
        idx, seed = int(fx["user_index"][0]), int(fx["balance_seed"][0])
        bal = np.random.default_rng(seed).integers(0, 1 << 40, size=(1 << levels, nc), dtype=np.uint64)[idx]
        entry = M.Entry("user_%d" % idx, [int(x) for x in bal])
        pr = {"entry": _E(entry.preimage()), "sibling_leaf_node_hash_preimage": [um(v) for v in fx["sibling_leaf_preimage"]],
              "sibling_middle_node_hash_preimages": [[um(v) for v in pre] for pre in fx["sibling_middle_preimages"]], "path_indices": [int(b) for b in fx["path_indices"]]}
        inst, got = _product_witness(pr, levels, nc, k)
        assert got == {(int(c), int(r)): um(v) for (c, r), v in zip(fx["advice_cells"], fx["advice_values"])}
        assert inst == [um(v) for v in fx["instances"]]


def test_product_witness_rejects_what_the_circuit_cannot_satisfy():
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200._lib import SummaB200Error
    from oracle import mst as M
    t = M.MerkleSumTree([M.Entry(f"u{i}", [(1 << 63) + i, 1]) for i in range(4)])   # sums leave the 8-byte range at level 1
    with pytest.raises(SummaB200Error):
        _product_witness(t.generate_proof(0), 2, 2, 11)
    ok = M.MerkleSumTree.from_csv(os.path.join(ROOT, "tests", "golden", "entry_16.csv"))
    with pytest.raises(SummaB200Error):
        _product_witness(ok.generate_proof(0), 4, 2, 10)   # LEVELS = 4 needs k = 11 (circuits/tests.rs:23): NotEnoughRowsAvailable


def test_evaluate_h_generated_source_compiles_offline(tmp_path):
    """the CUDA source csrc/expr_jit.cu generates for the reference circuit's quotient program (what NVRTC compiles on the GPU box) compiles for
    sm_100a with nvcc here, without spills; one non-inlined Montgomery product, straight-line body"""
    from circuits_halo2_b200 import _lib
    L = _lib.lib()
    for name in ("mst_inclusion_cs.json", "mst_inclusion_cs_n8.json"):
        cs = open(os.path.join(ROOT, "tests", "golden", name)).read().encode()
        n = ctypes.c_size_t()
        assert L.sb_test_h_jit_source(cs, None, ctypes.c_size_t(0), ctypes.byref(n)) == 0
        buf = ctypes.create_string_buffer(n.value + 1)
        assert L.sb_test_h_jit_source(cs, buf, ctypes.c_size_t(n.value), ctypes.byref(n)) == 0
        src = tmp_path / (name + ".cu")
        src.write_bytes(buf.raw[: n.value])
        out = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-cubin", "-Xptxas", "-v", "-o", str(tmp_path / "h.cubin"), str(src)],
                             capture_output=True, text=True)
        assert out.returncode == 0, out.stderr[-2000:]
        assert "0 bytes spill stores" in out.stderr and "sb_h_jit" in out.stderr
        body = buf.raw[: n.value].decode()
        assert body.count("jmul(") >= 100 and "for (" not in body.split("sb_h_jit")[-1]


def test_msm_host_tail_bucket_tree_records_and_residue_shards():
    """host side of a table MSM (csrc/host_g1.cpp: Horner fold over the bucket tree's bit sums, the running-sum pre-level's shift, the
    bucket-residue rescaling, normalisation by the binary-Euclid inversion), on the CPU against the oracle's group law"""
    from circuits_halo2_b200 import _lib
    L = _lib.lib()
    rnd = random.Random(23)

    def xyzz(p):   # affine point (or None) as a 128-byte XYZZ record, zz = zzz = 1 (identity: zeros)
        if p is None:
            return bytes(128)
        one = B.fq_to_mont_bytes(1) if hasattr(B, "fq_to_mont_bytes") else ((1 << 256) % B.Q).to_bytes(32, "little")
        return B.g1_to_mont_bytes(p) + one + one

    for trial in range(12):
        n_bits = rnd.choice([0, 3, 8, 16])
        shift = rnd.choice([0, 0, 5])
        log_mod = rnd.choice([0, 1, 3])
        res = rnd.randrange(1 << log_mod)
        x_slot = 17 if shift else 0
        pts = [B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R)) if rnd.random() < 0.85 else None for _ in range(18)]
        fin = np.frombuffer(b"".join(xyzz(p) for p in pts), dtype=np.uint8).copy()
        exp = None
        for b in range(n_bits):
            if pts[1 + b] is not None:
                exp = B.g1_add(exp, B.g1_mul(pts[1 + b], 1 << (b + shift)))
        exp = B.g1_add(exp, pts[x_slot])
        if log_mod:
            exp = B.g1_mul(exp, 1 << log_mod) if exp is not None else None
            k = (1 << log_mod) - res - 1
            if k and pts[0] is not None:
                exp = B.g1_add(exp, B.g1_neg(B.g1_mul(pts[0], k)))
        out = np.zeros(64, dtype=np.uint8)
        assert L.sb_test_msm_host_tail(_p(fin), n_bits, shift, x_slot, log_mod, res, _p(out)) == 0
        assert B.g1_from_mont_bytes(out.tobytes()) == exp, (trial, n_bits, shift, log_mod, res)


def test_row_sharded_grand_products_model():
    """control logic of the row-sharded grand products of a sharded proof (csrc/prover.cu, gp_rows), over integers mod r: every rank scans its row
    range from 1, the slice totals and the last rank's un-chained boundary values meet in one host exchange, every rank scales its slice by the
    product of the slices before it (and, for permutation sets, by the chained boundary values of the sets before) -- equal to the single-GPU
    columns: running products chained from set to set at row n - bf - 1."""
    rnd = random.Random(31)
    R = B.R
    for world, n, n_sets, n_lk, bf in [(2, 64, 2, 1, 5), (4, 64, 3, 0, 5), (8, 256, 1, 2, 7), (4, 128, 2, 2, 3)]:
        n_z = n_sets + n_lk
        ratio = [[rnd.randrange(1, R) for _ in range(n)] for _ in range(n_z)]
        b_row = n - bf - 1

        def scan(a, init=1):   # fr_running_product: z[i] = init * prod_{j < i} a[j]
            z, cur = [], init
            for x in a:
                z.append(cur)
                cur = cur * x % R
            return z

        # single GPU
        ref = [scan(ratio[z]) for z in range(n_z)]
        ref_unchained = [list(c) for c in ref]         # the library reads every boundary value first and scales afterwards
        carry = 1
        for s in range(1, n_sets):
            carry = carry * ref_unchained[s - 1][b_row] % R
            ref[s] = [v * carry % R for v in ref_unchained[s]]
        # sharded
        cnt = n // world
        assert bf + 1 < cnt
        local = [[scan(ratio[z][r * cnt:(r + 1) * cnt]) for z in range(n_z)] for r in range(world)]
        rec = [[local[r][z][cnt - 1] * ratio[z][(r + 1) * cnt - 1] % R for z in range(n_z)] +
               [local[r][z][b_row - r * cnt] if r == world - 1 else 0 for z in range(n_z)] for r in range(world)]   # the allgather_host records
        out = [[None] * n for _ in range(n_z)]
        for r in range(world):
            carry = 1
            for z in range(n_z):
                before = before_last = 1
                for q in range(world - 1):
                    if q < r:
                        before = before * rec[q][z] % R
                    before_last = before_last * rec[q][z] % R
                perm = z < n_sets
                scale = before * carry % R if perm else before
                for i in range(cnt):
                    out[z][r * cnt + i] = local[r][z][i] * scale % R
                if perm:
                    carry = carry * (before_last * rec[world - 1][n_z + z] % R) % R
        assert out == ref, (world, n, n_sets, n_lk)
