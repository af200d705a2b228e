"""CPU ORACLE (test infrastructure, NOT the product) -- BN254 arithmetic in Python big ints.

This module is a *checker*: it restates, in plain Python integers, the arithmetic of the
un-vendored crates the reference calls (halo2curves 0.1.0 `bn256::{Fr,Fq,G1Affine,G1}`,
halo2_proofs 0.2.0 @ summa-dev/halo2#8386d6e `arithmetic::{best_multiexp,best_fft}`,
`poly::EvaluationDomain`, `poly::kzg::commitment::ParamsKZG`).  Reference call sites:
`zk_prover/src/circuits/utils.rs:10-26,55,64,70,75,76,94-102`.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline leg may import it.
Pinned by: `tests/test_oracle_golden.py` (verifier-contract constants, SRS file, MSM KAT).
Byte conventions follow halo2curves: a field element is 4 x u64 little-endian limbs in
Montgomery form (R = 2^256); `G1Affine` = x || y (64 B), identity = (0, 0).
"""
from __future__ import annotations

import struct
from typing import Iterable, List, Optional, Sequence, Tuple

# --- moduli (contracts/src/InclusionVerifier.sol:209-210) ------------------------------------
Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583  # base field
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617  # scalar field
MONT = 1 << 256
MONT_INV_R = pow(MONT, -1, R)
MONT_INV_Q = pow(MONT, -1, Q)

# --- Fr constants (halo2curves bn256::Fr) -------------------------------
S = 28
GENERATOR = 7
ROOT_OF_UNITY = pow(GENERATOR, (R - 1) >> S, R)
DELTA = pow(GENERATOR, 1 << S, R)
ZETA = 0x30644E72E131A029048B6E193FD84104CC37A73FEC2BC5E9B8CA0B2D36636F23  # cube root of unity
CURVE_B = 3
G1_GEN = (1, 2)

Point = Optional[Tuple[int, int]]  # affine, None = identity


# --- byte <-> int (halo2curves in-memory layout) ---------------------------------------------
def fr_to_mont_bytes(x: int) -> bytes:
    return ((x * MONT) % R).to_bytes(32, "little")


def fr_from_mont_bytes(b: bytes) -> int:
    return (int.from_bytes(b, "little") * MONT_INV_R) % R


def fq_to_mont_bytes(x: int) -> bytes:
    return ((x * MONT) % Q).to_bytes(32, "little")


def fq_from_mont_bytes(b: bytes) -> int:
    return (int.from_bytes(b, "little") * MONT_INV_Q) % Q


def g1_to_mont_bytes(p: Point) -> bytes:
    if p is None:
        return b"\x00" * 64
    return fq_to_mont_bytes(p[0]) + fq_to_mont_bytes(p[1])


def g1_from_mont_bytes(b: bytes) -> Point:
    x = fq_from_mont_bytes(b[:32])
    y = fq_from_mont_bytes(b[32:64])
    if x == 0 and y == 0:
        return None
    return (x, y)


def frs_to_bytes(xs: Iterable[int]) -> bytes:
    return b"".join(fr_to_mont_bytes(x) for x in xs)


def frs_from_bytes(b: bytes) -> List[int]:
    return [fr_from_mont_bytes(b[i : i + 32]) for i in range(0, len(b), 32)]


def g1s_to_bytes(ps: Iterable[Point]) -> bytes:
    return b"".join(g1_to_mont_bytes(p) for p in ps)


def g1s_from_bytes(b: bytes) -> List[Point]:
    return [g1_from_mont_bytes(b[i : i + 64]) for i in range(0, len(b), 64)]


# --- G1 (y^2 = x^3 + 3) --------------------------------------------------------------------
def g1_is_on_curve(p: Point) -> bool:
    if p is None:
        return True
    x, y = p
    return (y * y - x * x * x - CURVE_B) % Q == 0


def g1_neg(p: Point) -> Point:
    if p is None:
        return None
    return (p[0], (-p[1]) % Q)


def g1_add(p: Point, q: Point) -> Point:
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if (y1 + y2) % Q == 0:
            return None
        lam = (3 * x1 * x1) * pow(2 * y1, -1, Q) % Q
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, Q) % Q
    x3 = (lam * lam - x1 - x2) % Q
    y3 = (lam * (x1 - x3) - y1) % Q
    return (x3, y3)


# Jacobian internals for speed (no inversion per add)
def _jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return P
    A = X * X % Q
    B = Y * Y % Q
    C = B * B % Q
    D = 2 * ((X + B) * (X + B) - A - C) % Q
    E = 3 * A % Q
    F = E * E % Q
    X3 = (F - 2 * D) % Q
    Y3 = (E * (D - X3) - 8 * C) % Q
    Z3 = 2 * Y * Z % Q
    return (X3, Y3, Z3)


def _jac_add(P, Qp):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Qp
    if Z1 == 0:
        return Qp
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % Q
    Z2Z2 = Z2 * Z2 % Q
    U1 = X1 * Z2Z2 % Q
    U2 = X2 * Z1Z1 % Q
    S1 = Y1 * Z2 * Z2Z2 % Q
    S2 = Y2 * Z1 * Z1Z1 % Q
    if U1 == U2:
        if S1 == S2:
            return _jac_double(P)
        return (0, 1, 0)
    H = (U2 - U1) % Q
    Rr = (S2 - S1) % Q
    HH = H * H % Q
    HHH = H * HH % Q
    V = U1 * HH % Q
    X3 = (Rr * Rr - HHH - 2 * V) % Q
    Y3 = (Rr * (V - X3) - S1 * HHH) % Q
    Z3 = Z1 * Z2 * H % Q
    return (X3, Y3, Z3)


def _to_jac(p: Point):
    return (0, 1, 0) if p is None else (p[0], p[1], 1)


def _from_jac(P) -> Point:
    X, Y, Z = P
    if Z == 0:
        return None
    zi = pow(Z, -1, Q)
    zi2 = zi * zi % Q
    return (X * zi2 % Q, Y * zi2 * zi % Q)


def g1_mul(p: Point, k: int) -> Point:
    k %= R
    acc = (0, 1, 0)
    base = _to_jac(p)
    while k:
        if k & 1:
            acc = _jac_add(acc, base)
        base = _jac_double(base)
        k >>= 1
    return _from_jac(acc)


def msm_naive(scalars: Sequence[int], bases: Sequence[Point]) -> Point:
    """sum_i scalars[i] * bases[i] by double-and-add (definition of `best_multiexp`'s result)."""
    acc = (0, 1, 0)
    for s, b in zip(scalars, bases):
        if s % R == 0 or b is None:
            continue
        acc = _jac_add(acc, _to_jac(g1_mul(b, s)))
    return _from_jac(acc)


def msm_pippenger(scalars: Sequence[int], bases: Sequence[Point], c: int = 8) -> Point:
    """Bucket method (the algorithm of halo2 `multiexp_serial`), fixed window c."""
    segments = 256 // c + 1
    acc = (0, 1, 0)
    for seg in reversed(range(segments)):
        for _ in range(c):
            acc = _jac_double(acc)
        buckets = [(0, 1, 0)] * ((1 << c) - 1)
        for s, b in zip(scalars, bases):
            d = (s >> (seg * c)) & ((1 << c) - 1)
            if d and b is not None:
                buckets[d - 1] = _jac_add(buckets[d - 1], _to_jac(b))
        run = (0, 1, 0)
        for bk in reversed(buckets):
            run = _jac_add(run, bk)
            acc = _jac_add(acc, run)
    return _from_jac(acc)


# --- NTT / EvaluationDomain (halo2_proofs::arithmetic::best_fft, poly::EvaluationDomain) --------
def omega_for(k: int) -> int:
    """omega = ROOT_OF_UNITY^(2^(S-k)) (EvaluationDomain::new)."""
    w = ROOT_OF_UNITY
    for _ in range(S - k):
        w = w * w % R
    return w


def bitrev(i: int, bits: int) -> int:
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def best_fft(a: List[int], omega: int, log_n: int) -> List[int]:
    """Natural-order in, natural-order out: out[k] = sum_j a[j] * omega^(j k)  (SURVEY A.3)."""
    n = 1 << log_n
    a = list(a)
    assert len(a) == n
    for i in range(n):
        j = bitrev(i, log_n)
        if i < j:
            a[i], a[j] = a[j], a[i]
    tw = [1] * (n // 2 if n > 1 else 1)
    for i in range(1, n // 2):
        tw[i] = tw[i - 1] * omega % R
    m = 1
    while m < n:
        step = n // (2 * m)
        for s in range(0, n, 2 * m):
            for j in range(m):
                t = a[s + j + m] * tw[j * step] % R
                u = a[s + j]
                a[s + j] = (u + t) % R
                a[s + j + m] = (u - t) % R
        m *= 2
    return a


def dft_naive(a: Sequence[int], omega: int) -> List[int]:
    n = len(a)
    return [sum(a[j] * pow(omega, j * k, R) for j in range(n)) % R for k in range(n)]


class EvaluationDomain:
    """Restates halo2 `EvaluationDomain::new(j, k)` and its transforms."""

    def __init__(self, j: int, k: int):
        self.k = k
        self.n = 1 << k
        qd = j - 1
        ext = k
        while (1 << ext) < self.n * qd:
            ext += 1
        self.quotient_poly_degree = qd
        self.extended_k = ext
        self.omega = omega_for(k)
        self.omega_inv = pow(self.omega, -1, R)
        self.extended_omega = omega_for(ext)
        self.extended_omega_inv = pow(self.extended_omega, -1, R)
        self.g_coset = ZETA
        self.g_coset_inv = ZETA * ZETA % R
        self.ifft_divisor = pow(self.n, -1, R)
        self.extended_ifft_divisor = pow(1 << ext, -1, R)
        # t(X) = X^n - 1 evaluated over the extended coset has 2^(ext-k) distinct values
        self.t_evaluations = []
        cur = pow(self.g_coset, self.n, R)
        orig = cur
        step = pow(self.extended_omega, self.n, R)
        for _ in range(1 << (ext - k)):
            self.t_evaluations.append((cur - 1) % R)
            cur = cur * step % R
        assert cur == orig
        self.t_inv = [pow(t, -1, R) for t in self.t_evaluations]

    def lagrange_to_coeff(self, a: List[int]) -> List[int]:
        out = best_fft(a, self.omega_inv, self.k)
        return [x * self.ifft_divisor % R for x in out]

    def coeff_to_lagrange(self, a: List[int]) -> List[int]:
        return best_fft(a, self.omega, self.k)

    def coeff_to_extended(self, a: List[int]) -> List[int]:
        cos = [1, self.g_coset, self.g_coset_inv]  # zeta^(i mod 3)
        b = [x * cos[i % 3] % R for i, x in enumerate(a)]
        b += [0] * ((1 << self.extended_k) - len(b))
        return best_fft(b, self.extended_omega, self.extended_k)

    def extended_to_coeff(self, a: List[int]) -> List[int]:
        b = best_fft(a, self.extended_omega_inv, self.extended_k)
        cos = [1, self.g_coset_inv, self.g_coset]  # zeta^-(i mod 3)
        b = [x * self.extended_ifft_divisor % R * cos[i % 3] % R for i, x in enumerate(b)]
        return b[: self.n * self.quotient_poly_degree]

    def divide_by_vanishing_poly(self, a: List[int]) -> List[int]:
        m = len(self.t_inv)
        return [x * self.t_inv[i % m] % R for i, x in enumerate(a)]

    def rotate_omega(self, x: int, rot: int) -> int:
        return x * pow(self.omega, rot, R) % R


# --- KZG params (poly::kzg::commitment::ParamsKZG) -------------------------------------------
class ParamsKZG:
    """`ParamsKZG::read` raw layout:
    u32 k || n x G1 (monomial) || n x G1 (Lagrange) || G2 || s.G2, raw Montgomery LE limbs."""

    def __init__(self, k: int, g_bytes: bytes, g_lagrange_bytes: bytes, tail: bytes = b""):
        self.k = k
        self.n = 1 << k
        self.g_bytes = g_bytes
        self.g_lagrange_bytes = g_lagrange_bytes
        self.tail = tail

    @classmethod
    def read(cls, path: str) -> "ParamsKZG":
        with open(path, "rb") as f:
            data = f.read()
        (k,) = struct.unpack("<I", data[:4])
        n = 1 << k
        g = data[4 : 4 + 64 * n]
        gl = data[4 + 64 * n : 4 + 128 * n]
        return cls(k, g, gl, data[4 + 128 * n :])

    @classmethod
    def setup_unsafe(cls, k: int, tau: int) -> "ParamsKZG":
        """`ParamsKZG::setup` with a *known* tau (UNSAFE test SRS): g[i] = tau^i G,
        g_lagrange[i] = L_i(tau) G.  Small k only (python speed)."""
        n = 1 << k
        g = []
        t = 1
        for _ in range(n):
            g.append(g1_mul(G1_GEN, t))
            t = t * tau % R
        gl = [g1_mul(G1_GEN, l) for l in lagrange_at(k, tau)]
        return cls(k, g1s_to_bytes(g), g1s_to_bytes(gl))

    def g(self) -> List[Point]:
        return g1s_from_bytes(self.g_bytes)

    def g_lagrange(self) -> List[Point]:
        return g1s_from_bytes(self.g_lagrange_bytes)


def lagrange_at(k: int, tau: int) -> List[int]:
    """L_i(tau) for the size-2^k domain: L_i(tau) = omega^i (tau^n - 1) / (n (tau - omega^i))."""
    n = 1 << k
    w = omega_for(k)
    tn = (pow(tau, n, R) - 1) % R
    ninv = pow(n, -1, R)
    out = []
    wi = 1
    for _ in range(n):
        out.append(wi * tn % R * ninv % R * pow((tau - wi) % R, -1, R) % R)
        wi = wi * w % R
    return out
