"""GPU parity tests (run on the B200 box: python -m pytest tests -m gpu).

Every test drives the CUDA path through the C ABI (ctypes -> libsumma_b200.so) and compares it
bit-for-bit with the CPU oracle on the same seeded inputs; large sizes use size-independent
properties (inverse round trip, linearity, split-and-sum).  Tolerance: none -- integer work is exact."""
import json
import os

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import cpu

pytestmark = pytest.mark.gpu


def fr_bytes(x):
    return np.frombuffer(B.fr_to_mont_bytes(x), dtype=np.uint64).copy()


# ------------------------------------------------------------------ field arithmetic (K1)
@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_vector_ops(ctx, field):
    import ctypes
    from circuits_halo2_b200 import _lib
    from circuits_halo2_b200.context import ptr
    mod = B.R if field == "fr" else B.Q
    n = 50000
    a = cpu.random_fr(n, 1)
    b = cpu.random_fr(n, 2)
    edge = [0, 1, mod - 1, mod - 2, (1 << 256) % mod, (1 << 254) % mod]
    for i, e in enumerate(edge):
        a[i] = np.frombuffer(e.to_bytes(32, "little"), dtype=np.uint64)
        b[len(edge) - 1 - i] = np.frombuffer(e.to_bytes(32, "little"), dtype=np.uint64)
    fn = getattr(_lib.lib(), f"sb_{field}_vec_op")
    ref = {0: getattr(cpu, f"{field}_mul"), 1: getattr(cpu, f"{field}_add"), 2: getattr(cpu, f"{field}_sub")}
    for op in (0, 1, 2):
        out = np.empty_like(a)
        _lib.check(fn(ctx.handle, ctypes.c_int32(op), ptr(a), ptr(b), ptr(out), ctypes.c_size_t(n)), "vec_op")
        assert (out.reshape(-1) == ref[op](a, b)).all(), (field, op)


# ------------------------------------------------------------------ NTT (K3)
@pytest.mark.parametrize("log_n", list(range(0, 15)) + [16, 17, 18])
def test_best_fft_matches_oracle(ctx, log_n):
    import circuits_halo2_b200 as sb
    a = cpu.random_fr(1 << log_n, 100 + log_n)
    w = fr_bytes(B.omega_for(log_n))
    got = sb.best_fft(a.copy(), w, log_n, ctx)
    ref = cpu.best_fft(a, w, log_n, threads=8)
    assert (got.reshape(-1) == ref).all()


def test_best_fft_inverse_omega_and_edge_inputs(ctx):
    import circuits_halo2_b200 as sb
    log_n = 12
    n = 1 << log_n
    w_inv = fr_bytes(pow(B.omega_for(log_n), -1, B.R))
    for name, a in (("zeros", np.zeros((n, 4), dtype=np.uint64)),
                    ("delta", np.concatenate([fr_bytes(1).reshape(1, 4), np.zeros((n - 1, 4), dtype=np.uint64)])),
                    ("max", np.tile(np.frombuffer((B.R - 1).to_bytes(32, "little"), dtype=np.uint64), (n, 1)))):
        got = sb.best_fft(a.copy(), w_inv, log_n, ctx)
        assert (got.reshape(-1) == cpu.best_fft(a, w_inv, log_n, threads=4)).all(), name


def test_best_fft_rejects_length_mismatch(ctx):
    import circuits_halo2_b200 as sb
    with pytest.raises(AssertionError):
        sb.best_fft(cpu.random_fr(100, 1), fr_bytes(B.omega_for(7)), 7, ctx)


def oracle_eval_poly(a, x):
    """sum_j a_j x^j with vectorised C-oracle field ops (powers by doubling, pairwise tree sum)."""
    n = a.shape[0]
    powers = fr_bytes(1).reshape(1, 4)
    xm = x
    while powers.shape[0] < n:
        m = powers.shape[0]
        nxt = cpu.fr_mul(powers.reshape(-1), np.tile(fr_bytes(xm), m)).reshape(m, 4)
        powers = np.concatenate([powers, nxt])
        xm = xm * xm % B.R
    acc = cpu.fr_mul(a.reshape(-1), powers.reshape(-1)).reshape(n, 4)
    while acc.shape[0] > 1:
        acc = cpu.fr_add(np.ascontiguousarray(acc[0::2]).reshape(-1), np.ascontiguousarray(acc[1::2]).reshape(-1)).reshape(-1, 4)
    return B.fr_from_mont_bytes(acc[0].tobytes())


@pytest.mark.parametrize("log_n", [20, 22, 24])
def test_best_fft_large_roundtrip_and_spot_check(ctx, log_n):
    """BASELINE sizes: two outputs checked against the definition X[k] = sum_j a_j w^(jk), and
    forward-then-inverse == n * identity."""
    import circuits_halo2_b200 as sb
    n = 1 << log_n
    a = cpu.random_fr(n, 7 + log_n)
    w = B.omega_for(log_n)
    f = sb.best_fft(a.copy(), fr_bytes(w), log_n, ctx)
    for k in (1, (123457 * 7919) % n):
        assert B.fr_from_mont_bytes(f[k].tobytes()) == oracle_eval_poly(a, pow(w, k, B.R)), k
    back = sb.best_fft(f, fr_bytes(pow(w, -1, B.R)), log_n, ctx)
    expect = cpu.fr_mul(a.reshape(-1), np.tile(fr_bytes(n), n))
    assert (back.reshape(-1) == expect).all()


# ------------------------------------------------------------------ EvaluationDomain (K4)
@pytest.mark.parametrize("k", [4, 11, 13])
def test_evaluation_domain_matches_oracle(ctx, k):
    import circuits_halo2_b200 as sb
    d = sb.EvaluationDomain(6, k, ctx)
    od = cpu.Domain(6, k, threads=8)
    assert d.extended_k() == k + 3
    a = cpu.random_fr(1 << k, 40 + k)
    coeff = d.lagrange_to_coeff(a)
    assert (coeff.reshape(-1) == od.lagrange_to_coeff(a.reshape(-1))).all()
    assert (d.coeff_to_lagrange(coeff).reshape(-1) == a.reshape(-1)).all()
    ext = d.coeff_to_extended(coeff)
    assert (ext.reshape(-1) == od.coeff_to_extended(coeff.reshape(-1))).all()
    div = d.divide_by_vanishing_poly(ext)
    assert (div.reshape(-1) == od.divide_by_vanishing_poly(ext.reshape(-1))).all()
    back = d.extended_to_coeff(ext)
    assert back.shape[0] == 5 << k
    assert (back.reshape(-1) == od.extended_to_coeff(ext.reshape(-1))).all()
    assert (back[: 1 << k] == coeff).all() and not back[1 << k:].any()


# ------------------------------------------------------------------ MSM (K2)
def scalar_distribution(kind, n, seed):
    a = cpu.random_fr(n, seed)
    rng = np.random.default_rng(seed + 1)
    if kind == "Z":    # 99 % zeros (advice-like)
        a[rng.random(n) < 0.99] = 0
    elif kind == "C":  # one repeated value on 90 % of the rows (Z-polynomial-like)
        a[rng.random(n) < 0.9] = a[0]
    elif kind == "S":  # 8-bit values (table-like), Montgomery form
        small = rng.integers(0, 256, size=n)
        table = np.concatenate([fr_bytes(int(v)).reshape(1, 4) for v in range(256)])
        a = table[small]
    elif kind == "E":  # edge values
        for i, v in enumerate([0, 1, B.R - 1, B.R - 2, (1 << 253), (1 << 128) - 1]):
            if i < n:
                a[i] = fr_bytes(v)
    return np.ascontiguousarray(a)


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 255, 1000, 4096, 1 << 14])
@pytest.mark.parametrize("kind", ["U", "Z", "C", "S", "E"])
def test_best_multiexp_matches_oracle(ctx, n, kind):
    import circuits_halo2_b200 as sb
    bases = cpu.gen_bases(n, seed=n, threads=4)
    sc = scalar_distribution(kind, n, 1000 + n)
    got = sb.best_multiexp(sc, bases, ctx)
    ref = cpu.best_multiexp(sc, bases, threads=8)
    if not ref.any():
        assert not got[8:].any()  # identity: z == 0
    else:
        assert (got[:8] == ref).all()
        assert B.fq_from_mont_bytes(got[8:].tobytes()) == 1


def test_best_multiexp_degenerate_bases(ctx):
    """identity bases, repeated bases (P + P inside a bucket), P and -P (cancellation)."""
    import circuits_halo2_b200 as sb
    n = 512
    bases = cpu.gen_bases(n, seed=77, threads=2)
    bases[5] = 0                      # identity point
    bases[10] = bases[11]             # duplicates
    neg = B.g1_neg(B.g1_from_mont_bytes(bases[20].tobytes()))
    bases[21] = np.frombuffer(B.g1_to_mont_bytes(neg), dtype=np.uint64)
    sc = cpu.random_fr(n, 78)
    sc[10] = sc[11]                   # same scalar on the duplicate -> same bucket, doubling path
    sc[20] = sc[21]                   # P and -P in the same bucket -> identity partial
    sc[100:200] = sc[100]             # a hot bucket
    ref = cpu.best_multiexp(sc, bases, threads=4)
    got = sb.best_multiexp(sc, bases, ctx)
    assert (got[:8] == ref).all()
    # all scalars equal on all-equal bases: n * s * P
    same_b = np.tile(bases[0], (n, 1))
    same_s = np.tile(sc[0], (n, 1))
    ref = cpu.best_multiexp(same_s, same_b, threads=4)
    assert (sb.best_multiexp(same_s, same_b, ctx)[:8] == ref).all()
    # empty input -> identity
    assert not sb.best_multiexp(np.zeros((0, 4), dtype=np.uint64), np.zeros((0, 8), dtype=np.uint64), ctx)[8:].any()
    with pytest.raises(AssertionError):
        sb.best_multiexp(sc[:5], bases[:6], ctx)


def test_commit_lagrange_golden_fixed_comm_4(ctx, golden_dir):
    """KAT from the reference's verifier contract (.sol:246-247) on the reference's own SRS file."""
    import circuits_halo2_b200 as sb
    vk = json.load(open(os.path.join(golden_dir, "verifier_constants.json")))
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    assert params.k() == 11
    table = np.zeros((params.n, 4), dtype=np.uint64)
    for i in range(256):
        table[i] = fr_bytes(i)
    got = params.commit_lagrange(table)
    assert B.g1_from_mont_bytes(got.tobytes()) == (int(vk["fixed_comms[4].x"], 16), int(vk["fixed_comms[4].y"], 16))
    # commit / commit_lagrange agree through the domain: commit(lagrange_to_coeff(v)) == commit_lagrange(v)
    d = sb.EvaluationDomain(6, 11, ctx)
    v = cpu.random_fr(params.n, 5)
    assert (params.commit(d.lagrange_to_coeff(v)) == params.commit_lagrange(v)).all()
    ones = np.tile(fr_bytes(1), (params.n, 1))
    assert (params.commit_lagrange(ones) == params.g[0]).all()  # sum of Lagrange bases = g[0]


@pytest.mark.parametrize("log_n,kind", [(18, "U"), (20, "U"), (20, "C"), (20, "Z")])
def test_best_multiexp_large_split_property(ctx, log_n, kind):
    """BASELINE sizes: MSM(s, P) == MSM(s[:h], P[:h]) + MSM(s[h:], P[h:]) (different window shapes
    on both sides), and linearity MSM(s + s', P) == MSM(s, P) + MSM(s', P); 2^18 also vs the oracle."""
    import ctypes
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import _lib
    from circuits_halo2_b200.context import ptr
    n = 1 << log_n
    bases = cpu.gen_bases(n, seed=log_n, threads=8)
    sc = scalar_distribution(kind, n, 500 + log_n)
    full = sb.best_multiexp(sc, bases, ctx)[:8]
    h = n // 3
    parts = np.stack([sb.best_multiexp(sc[:h], bases[:h], ctx)[:8], sb.best_multiexp(sc[h:], bases[h:], ctx)[:8]])
    out = np.zeros(8, dtype=np.uint64)
    _lib.check(_lib.lib().sb_g1_sum_affine(ptr(parts), ctypes.c_size_t(2), ptr(out)), "sum")
    assert (out == full).all()
    assert (out == cpu.g1_add(parts[0], parts[1])).all()
    if log_n <= 18:
        assert (full == cpu.best_multiexp(sc, bases, threads=8)).all()
    sc2 = cpu.random_fr(n, 900 + log_n)
    ssum = cpu.fr_add(sc.reshape(-1), sc2.reshape(-1)).reshape(n, 4)
    lhs = sb.best_multiexp(ssum, bases, ctx)[:8]
    rhs = cpu.g1_add(full, sb.best_multiexp(sc2, bases, ctx)[:8])
    assert (lhs == rhs).all()


@pytest.mark.parametrize("k,window_bits", [(11, 0), (12, 13), (14, 0), (14, 16), (16, 20)])
def test_fixed_base_tables_give_identical_commitments(ctx, k, window_bits):
    """sb_srs_precompute (2^(c w) * P_i tables, one shared bucket set) must not change a single bit of commit / commit_lagrange:
    every scalar distribution, prefixes n < 2^k, identity and repeated bases in the SRS, all checked against the table-free path
    and (k <= 14) against the oracle."""
    import circuits_halo2_b200 as sb
    n = 1 << k
    g = cpu.gen_bases(n, seed=40 + k, threads=8)
    gl = cpu.gen_bases(n, seed=90 + k, threads=8)
    g[3] = 0
    g[7] = g[8]
    gl[0] = 0
    plain = sb.ParamsKZG(k, g, gl, ctx=ctx)
    tabbed = sb.ParamsKZG(k, g, gl, ctx=ctx).precompute(3, window_bits)
    for kind in ["U", "Z", "C", "S", "E"]:
        sc = scalar_distribution(kind, n, 7000 + k)
        sc[7] = sc[8]
        a, b = tabbed.commit(sc), plain.commit(sc)
        assert (a == b).all(), (kind, "commit")
        assert (tabbed.commit_lagrange(sc) == plain.commit_lagrange(sc)).all(), (kind, "commit_lagrange")
        if k <= 14 and kind in ("U", "E"):
            assert (a == cpu.best_multiexp(sc, g, threads=8)).all()
    for m in (1, 5, n // 2 + 3, n - 1):
        sc = cpu.random_fr(m, 7100 + m)
        assert (tabbed.commit(sc) == plain.commit(sc)).all(), m
    assert not tabbed.commit(np.zeros((n, 4), dtype=np.uint64)).any()


def test_fixed_base_tables_large(ctx):
    """2^20 bases, c = 20 (13 windows, 2^19 shared buckets): equal to the table-free MSM on uniform and constant-heavy scalars."""
    import ctypes
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import _lib
    k = 20
    params = sb.ParamsKZG.setup(k, 0x5A110000 + k, ctx, download=False)
    res = {}
    for kind in ["U", "C", "Z"]:
        sc = scalar_distribution(kind, 1 << k, 7200)
        res[kind] = (params.commit(sc), params.commit_lagrange(sc))
    params.precompute()
    for kind in ["U", "C", "Z"]:
        sc = scalar_distribution(kind, 1 << k, 7200)
        assert (params.commit(sc) == res[kind][0]).all() and (params.commit_lagrange(sc) == res[kind][1]).all(), kind


def test_params_downsize_matches_g_to_lagrange(ctx, golden_dir):
    """ParamsKZG::downsize on the reference's own SRS (hermez-raw-11): the first 2^k monomial bases are kept and the Lagrange bases
    equal the definition g_lagrange[i] = MSM(coefficients of L_i, g) computed by the oracle (k = 4), and commit_lagrange(v) ==
    commit(lagrange_to_coeff(v)) for the downsized params (k = 4, 7, 10); downsize(11) is the identity."""
    import circuits_halo2_b200 as sb
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    same = params.downsize(11)
    assert (same.g == params.g).all() and (same.g_lagrange == params.g_lagrange).all()
    for k in (4, 7, 10):
        small = params.downsize(k)
        n = 1 << k
        assert (small.g == params.g[:n]).all()
        dom = sb.EvaluationDomain(6, k, ctx)
        v = cpu.random_fr(n, 600 + k)
        assert (small.commit_lagrange(v) == small.commit(dom.lagrange_to_coeff(v))).all(), k
        if k == 4:
            for i in (0, 1, 5, 15):
                e = np.zeros((n, 4), dtype=np.uint64)
                e[i] = fr_bytes(1)
                li = dom.lagrange_to_coeff(e)  # coefficients of the i-th Lagrange polynomial
                assert (small.g_lagrange[i] == cpu.best_multiexp(np.ascontiguousarray(li.reshape(n, 4)), np.ascontiguousarray(params.g[:n]), threads=2)).all(), i
    with pytest.raises(AssertionError):
        params.downsize(12)
