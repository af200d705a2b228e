"""Pins the CPU oracle (oracle/bn254.py + oracle/halo2_cpu.c) on the reference's golden vectors
(G2 SRS file, G3 domain constants, G4 fixed_comms[4] MSM KAT) -- CPU only."""
import json
import os

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import cpu


@pytest.fixture(scope="module")
def vk(golden_dir):
    return json.load(open(os.path.join(golden_dir, "verifier_constants.json")))


@pytest.fixture(scope="module")
def srs(golden_dir):
    return B.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"))


def test_field_constants():
    assert pow(B.ROOT_OF_UNITY, 1 << 28, B.R) == 1 and pow(B.ROOT_OF_UNITY, 1 << 27, B.R) != 1
    assert B.ROOT_OF_UNITY == 0x03DDB9F5166D18B798865EA93DD31F743215CF6DD39329C8D34F1ED960C37C9C
    assert pow(B.ZETA, 3, B.R) == 1 and B.ZETA != 1


def test_domain_constants_match_verifier_contract(vk):
    # contracts/src/InclusionVerifier.sol:218-222
    k = int(vk["k"], 16)
    assert k == 11
    d = B.EvaluationDomain(6, k)
    assert d.ifft_divisor == int(vk["n_inv"], 16)
    assert d.omega == int(vk["omega"], 16)
    assert d.omega_inv == int(vk["omega_inv"], 16)
    assert pow(d.omega_inv, 6, B.R) == int(vk["omega_inv_to_l"], 16)  # blinding_factors = 5
    assert d.extended_k == k + 3 and len(d.t_evaluations) == 8
    assert B.DELTA == int(vk["delta"], 16)  # .sol:498


def test_srs_file_layout(srs, vk, golden_dir):
    assert os.path.getsize(os.path.join(golden_dir, "hermez-raw-11")) == 4 + 2 * 2048 * 64 + 2 * 128
    assert srs.k == 11
    g = srs.g()
    assert g[0] == (int(vk["g1_x"], 16), int(vk["g1_y"], 16)) == (1, 2)
    gl = srs.g_lagrange()
    for p in (g[1], g[2047], gl[0], gl[1], gl[2047]):
        assert B.g1_is_on_curve(p)
    # sum of the Lagrange basis commitments = commitment to the constant 1 = g[0]
    acc = None
    for p in gl:
        acc = B.g1_add(acc, p)
    assert acc == g[0]


def test_msm_kat_fixed_comm_4(srs, vk):
    """fixed_comms[4] (.sol:246-247) = commit_lagrange of the 0..255 range table column."""
    expect = (int(vk["fixed_comms[4].x"], 16), int(vk["fixed_comms[4].y"], 16))
    scalars = [i if i < 256 else 0 for i in range(srs.n)]
    gl = srs.g_lagrange()
    assert B.msm_pippenger(scalars[:256], gl[:256], 8) == expect
    out = cpu.best_multiexp(B.frs_to_bytes(scalars), srs.g_lagrange_bytes, threads=4)
    assert B.g1_from_mont_bytes(out.tobytes()) == expect


def test_c_oracle_field_ops_match_python():
    import random
    rnd = random.Random(11)
    for mod, mul, add, sub, to_b in ((B.R, cpu.fr_mul, cpu.fr_add, cpu.fr_sub, B.fr_to_mont_bytes),
                                     (B.Q, cpu.fq_mul, cpu.fq_add, cpu.fq_sub, B.fq_to_mont_bytes)):
        xs = [0, 1, mod - 1] + [rnd.randrange(mod) for _ in range(200)]
        ys = [mod - 1, 0, mod - 1] + [rnd.randrange(mod) for _ in range(200)]
        a = np.frombuffer(b"".join(to_b(x) for x in xs), dtype=np.uint64)
        b = np.frombuffer(b"".join(to_b(y) for y in ys), dtype=np.uint64)
        inv = pow(1 << 256, -1, mod)
        dec = lambda arr: [int.from_bytes(arr[4 * i:4 * i + 4].tobytes(), "little") * inv % mod for i in range(len(xs))]
        assert dec(mul(a, b)) == [x * y % mod for x, y in zip(xs, ys)]
        assert dec(add(a, b)) == [(x + y) % mod for x, y in zip(xs, ys)]
        assert dec(sub(a, b)) == [(x - y) % mod for x, y in zip(xs, ys)]


@pytest.mark.parametrize("k", [1, 2, 5, 9])
def test_c_oracle_fft_matches_definition(k):
    import random
    rnd = random.Random(k)
    w = B.omega_for(k)
    v = [rnd.randrange(B.R) for _ in range(1 << k)]
    ref = B.dft_naive(v, w) if k <= 5 else B.best_fft(v, w, k)
    for threads in (1, 4):
        out = cpu.best_fft(B.frs_to_bytes(v), B.fr_to_mont_bytes(w), k, threads)
        assert B.frs_from_bytes(out.tobytes()) == ref


def test_c_oracle_msm_matches_naive():
    import random
    rnd = random.Random(5)
    n = 70
    pts = [B.g1_mul(B.G1_GEN, rnd.randrange(1, B.R)) for _ in range(n)]
    sc = [rnd.randrange(B.R) for _ in range(n)]
    sc[3] = 0
    sc[4] = B.R - 1
    pts[5] = None
    pts[7] = pts[8]
    ref = B.msm_naive(sc, pts)
    for threads in (1, 3, 8):
        out = cpu.best_multiexp(B.frs_to_bytes(sc), B.g1s_to_bytes(pts), threads)
        assert B.g1_from_mont_bytes(out.tobytes()) == ref


def test_c_oracle_domain_roundtrip():
    d = cpu.Domain(6, 5)
    a = cpu.random_fr(32, 3).reshape(-1)
    coeff = d.lagrange_to_coeff(a)
    assert (d.coeff_to_lagrange(coeff) == a).all()
    ext = d.coeff_to_extended(coeff)
    back = d.extended_to_coeff(ext)
    assert (back[: 32 * 4] == coeff).all() and not back[32 * 4:].any()
    # python twin agrees
    pd = B.EvaluationDomain(6, 5)
    assert B.frs_from_bytes(ext.tobytes()) == pd.coeff_to_extended(B.frs_from_bytes(coeff.tobytes()))


def test_gen_bases_are_an_arithmetic_progression():
    b = cpu.gen_bases(300, seed=3, threads=3)
    pts = B.g1s_from_bytes(b.tobytes())
    assert all(B.g1_is_on_curve(p) for p in pts[:4] + pts[-4:])
    d = B.g1_add(pts[1], B.g1_neg(pts[0]))
    for i in (1, 99, 100, 101, 199, 200, 298):
        assert B.g1_add(pts[i], d) == pts[i + 1]
