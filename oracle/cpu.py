"""ctypes front-end of oracle/halo2_cpu.c (CPU ORACLE -- test infrastructure, NOT the product).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
All arrays are numpy uint64 views of halo2curves' memory layout (Montgomery, 4 x u64 LE limbs).
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libhalo2_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "halo2_cpu.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.oracle_init()
    return _lib


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


def _u64(buf) -> np.ndarray:
    if isinstance(buf, (bytes, bytearray)):
        return np.frombuffer(bytes(buf), dtype=np.uint64).copy()
    return np.ascontiguousarray(buf).view(np.uint64).reshape(-1)


def _binop(name, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(out), _p(a), _p(b), ctypes.c_size_t(a.size // 4))
    return out


def fr_mul(a, b): return _binop("oracle_fr_mul", a, b)
def fr_add(a, b): return _binop("oracle_fr_add", a, b)
def fr_sub(a, b): return _binop("oracle_fr_sub", a, b)
def fq_mul(a, b): return _binop("oracle_fq_mul", a, b)
def fq_add(a, b): return _binop("oracle_fq_add", a, b)
def fq_sub(a, b): return _binop("oracle_fq_sub", a, b)


def _unop(name, a):
    a = _u64(a)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(out), _p(a), ctypes.c_size_t(a.size // 4))
    return out


def fr_inv(a): return _unop("oracle_fr_inv", a)
def fr_from_canon(a): return _unop("oracle_fr_from_canon", a)
def fr_to_canon(a): return _unop("oracle_fr_to_canon", a)
def fq_from_canon(a): return _unop("oracle_fq_from_canon", a)
def fq_to_canon(a): return _unop("oracle_fq_to_canon", a)


def best_multiexp(coeffs, bases, threads: int = 1) -> np.ndarray:
    """halo2 `best_multiexp`; returns the affine result (8 x u64: x || y, Montgomery)."""
    c, b = _u64(coeffs), _u64(bases)
    n = c.size // 4
    assert b.size // 8 >= n
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_best_multiexp(_p(out), _p(c), _p(b), ctypes.c_size_t(n), ctypes.c_int(threads))
    return out


def best_fft(a, omega, log_n: int, threads: int = 1) -> np.ndarray:
    """halo2 `best_fft` (natural in / natural out). Returns a new array."""
    a = _u64(a).copy()
    w = _u64(omega)
    assert a.size == 4 << log_n
    lib().oracle_best_fft(_p(a), _p(w), ctypes.c_uint32(log_n), ctypes.c_int(threads))
    return a


def fr_scale(a, s) -> np.ndarray:
    a = _u64(a).copy()
    lib().oracle_fr_scale(_p(a), _p(_u64(s)), ctypes.c_size_t(a.size // 4))
    return a


def fr_scale_pattern(a, pat) -> np.ndarray:
    a = _u64(a).copy()
    pat = _u64(pat)
    lib().oracle_fr_scale_pattern(_p(a), _p(pat), ctypes.c_size_t(pat.size // 4), ctypes.c_size_t(a.size // 4))
    return a


def g1_mul(p, k_mont) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_g1_mul(_p(out), _p(_u64(p)), _p(_u64(k_mont)))
    return out


def g1_add(a, b) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_g1_add_affine(_p(out), _p(_u64(a)), _p(_u64(b)))
    return out


class Domain:
    """EvaluationDomain transforms on Montgomery byte arrays, C speed."""

    def __init__(self, j: int, k: int, threads: int = 1):
        from . import bn254 as B
        self.d = B.EvaluationDomain(j, k)
        self.threads = threads
        m = lambda x: np.frombuffer(B.fr_to_mont_bytes(x), dtype=np.uint64).copy()
        self.omega, self.omega_inv = m(self.d.omega), m(self.d.omega_inv)
        self.ext_omega, self.ext_omega_inv = m(self.d.extended_omega), m(self.d.extended_omega_inv)
        self.ifft_div, self.ext_ifft_div = m(self.d.ifft_divisor), m(self.d.extended_ifft_divisor)
        self.coset = np.concatenate([m(1), m(self.d.g_coset), m(self.d.g_coset_inv)])
        self.coset_inv = np.concatenate([m(1), m(self.d.g_coset_inv), m(self.d.g_coset)])
        self.t_inv = np.concatenate([m(t) for t in self.d.t_inv])

    def lagrange_to_coeff(self, a):
        return par_fr_scale(best_fft(a, self.omega_inv, self.d.k, self.threads), self.ifft_div)

    def coeff_to_lagrange(self, a):
        return best_fft(a, self.omega, self.d.k, self.threads)

    def coeff_to_extended(self, a):
        b = par_fr_scale_pattern(a, self.coset)
        ext = np.zeros(4 << self.d.extended_k, dtype=np.uint64)
        ext[: b.size] = b
        return best_fft(ext, self.ext_omega, self.d.extended_k, self.threads)

    def extended_to_coeff(self, a):
        b = best_fft(a, self.ext_omega_inv, self.d.extended_k, self.threads)
        b = par_fr_scale(b, self.ext_ifft_div)
        b = par_fr_scale_pattern(b, self.coset_inv)
        return b[: 4 * self.d.n * self.d.quotient_poly_degree].copy()

    def divide_by_vanishing_poly(self, a):
        return par_fr_scale_pattern(a, self.t_inv)


def gen_bases(n: int, seed: int = 1, threads: int = 8) -> np.ndarray:
    """n distinct valid G1Affine points (s + i t) G, as an (n, 8) uint64 Montgomery array."""
    from . import bn254 as B
    rng = np.random.default_rng(seed)
    s = int.from_bytes(rng.bytes(31), "little") % B.R
    t = int.from_bytes(rng.bytes(31), "little") % B.R or 1
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oracle_g1_gen_bases(_p(out), ctypes.c_size_t(n), _p(_u64(B.fr_to_mont_bytes(s))), _p(_u64(B.fr_to_mont_bytes(t))), ctypes.c_int(threads))
    return out


def random_fr(n: int, seed: int = 0) -> np.ndarray:
    """n field elements < 2^253 (< r), uniform on that range, as an (n, 4) uint64 array.
    Any value < r is a valid Montgomery residue, so no conversion is needed."""
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 1 << 63, size=(n, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)
    return a


def fr_batch_invert(a) -> np.ndarray:
    a = _u64(a).copy()
    lib().oracle_fr_batch_invert(_p(a), ctypes.c_size_t(a.size // 4))
    return a


def fr_running_product(a, init, n: int) -> np.ndarray:
    """z[0] = init, z[i] = z[i-1] * a[i-1] for i < n."""
    a = _u64(a)
    out = np.zeros(4 * n, dtype=np.uint64)
    lib().oracle_fr_running_product(_p(out), _p(a), _p(_u64(init)), ctypes.c_size_t(n))
    return out


def fr_eval_poly(coeffs, x) -> np.ndarray:
    c = _u64(coeffs)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_eval_poly(_p(out), _p(c), ctypes.c_size_t(c.size // 4), _p(_u64(x)))
    return out


def fr_kate_division(a, b) -> np.ndarray:
    a = _u64(a)
    n = a.size // 4
    q = np.zeros(4 * max(n - 1, 0), dtype=np.uint64)
    lib().oracle_fr_kate_division(_p(q), _p(a), ctypes.c_size_t(n), _p(_u64(b)))
    return q


def fr_axpby(a, s, b, t) -> np.ndarray:
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    lib().oracle_fr_axpby(_p(out), _p(a), _p(_u64(s)), _p(b), _p(_u64(t)), ctypes.c_size_t(a.size // 4))
    return out


# ---------------------------------------------------------------------------------------------------
# threaded forms (OpenMP inside halo2_cpu.c): what halo2 does with rayon's `parallelize`
def set_threads(t: int) -> None:
    lib().oracle_set_threads(ctypes.c_int(int(t)))


def get_threads() -> int:
    return int(lib().oracle_get_threads())


def _par_binop(name, a, b):
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    getattr(lib(), name)(_p(out), _p(a), _p(b), ctypes.c_size_t(a.size // 4))
    return out


def par_fr_mul(a, b): return _par_binop("oracle_par_fr_mul", a, b)
def par_fr_add(a, b): return _par_binop("oracle_par_fr_add", a, b)
def par_fr_sub(a, b): return _par_binop("oracle_par_fr_sub", a, b)


def par_fr_scale(a, s) -> np.ndarray:
    a = _u64(a).copy()
    lib().oracle_par_fr_scale(_p(a), _p(_u64(s)), ctypes.c_size_t(a.size // 4))
    return a


def par_fr_scale_pattern(a, pat) -> np.ndarray:
    a = _u64(a).copy()
    pat = _u64(pat)
    lib().oracle_par_fr_scale_pattern(_p(a), _p(pat), ctypes.c_size_t(pat.size // 4), ctypes.c_size_t(a.size // 4))
    return a


def par_fr_add_const(a, c) -> np.ndarray:
    a = _u64(a)
    out = np.empty_like(a)
    lib().oracle_par_fr_add_const(_p(out), _p(a), _p(_u64(c)), ctypes.c_size_t(a.size // 4))
    return out


def par_fr_axpby(a, s, b, t) -> np.ndarray:
    a, b = _u64(a), _u64(b)
    out = np.empty_like(a)
    lib().oracle_par_fr_axpby(_p(out), _p(a), _p(_u64(s)), _p(b), _p(_u64(t)), ctypes.c_size_t(a.size // 4))
    return out


def fr_powers(base, n: int) -> np.ndarray:
    """(n, 4): base^i in Montgomery form."""
    out = np.empty((n, 4), dtype=np.uint64)
    lib().oracle_fr_powers(_p(out), _p(_u64(base)), ctypes.c_size_t(n))
    return out


def par_fr_batch_invert(a) -> np.ndarray:
    a = _u64(a).copy()
    lib().oracle_par_fr_batch_invert(_p(a), ctypes.c_size_t(a.size // 4))
    return a


def par_fr_eval_poly(coeffs, x) -> np.ndarray:
    c = _u64(coeffs)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_par_fr_eval_poly(_p(out), _p(c), ctypes.c_size_t(c.size // 4), _p(_u64(x)))
    return out


def expr_eval(cols, code: np.ndarray, consts: np.ndarray, n_regs: int, out_reg: int, n_rows: int, rot_scale: int) -> np.ndarray:
    """Row-parallel register program over whole columns (halo2 `GraphEvaluator`): cols = list of (n_rows, 4) uint64 arrays,
    code = (n_ins, 4) int32 rows (op, dst, a, b), consts (m, 4) uint64.  Returns (n_rows, 4)."""
    keep = [np.ascontiguousarray(c).reshape(-1) for c in cols]
    for c in keep:
        assert c.size == 4 * n_rows
    ptrs = (ctypes.c_void_p * max(len(keep), 1))(*[c.ctypes.data for c in keep])
    code = np.ascontiguousarray(code, dtype=np.int32)
    consts = np.ascontiguousarray(consts, dtype=np.uint64).reshape(-1)
    out = np.empty((n_rows, 4), dtype=np.uint64)
    lib().oracle_expr_eval(_p(out), ctypes.c_size_t(n_rows), ptrs, _p(code), ctypes.c_size_t(code.shape[0]), _p(consts) if consts.size else None,
                           ctypes.c_int32(n_regs), ctypes.c_int32(out_reg), ctypes.c_int64(rot_scale))
    return out


def permute_expression_pair(inp, tab, usable: int):
    """halo2 lookup `permute_expression_pair` on Montgomery arrays; raises ValueError when an input is not in the table."""
    a, t = _u64(inp), _u64(tab)
    p_in = np.zeros((usable, 4), dtype=np.uint64)
    p_tab = np.zeros((usable, 4), dtype=np.uint64)
    rc = lib().oracle_permute_expression_pair(_p(p_in), _p(p_tab), _p(a), _p(t), ctypes.c_size_t(usable))
    if rc != 0:
        raise ValueError("lookup input value not in table (ConstraintSystemFailure)")
    return p_in, p_tab


def chacha_fr_fill(key_words, block0: int, n: int) -> np.ndarray:
    """n consecutive `Fr::random` draws of a ChaCha20Rng with this key, starting at keystream block `block0` (one draw = one block)."""
    key = np.array(list(key_words), dtype=np.uint32)
    out = np.empty((n, 4), dtype=np.uint64)
    lib().oracle_chacha_fr_fill(_p(out), _p(key), ctypes.c_uint64(block0), ctypes.c_size_t(n))
    return out


def keccak256(data: bytes) -> bytes:
    out = np.zeros(32, dtype=np.uint8)
    buf = np.frombuffer(data, dtype=np.uint8) if len(data) else np.zeros(1, dtype=np.uint8)
    lib().oracle_keccak256(_p(out), _p(np.ascontiguousarray(buf)), ctypes.c_size_t(len(data)))
    return out.tobytes()


_poseidon_ready = False


def _poseidon_init():
    global _poseidon_ready
    if _poseidon_ready:
        return
    import json
    from . import bn254 as B
    golden = os.path.join(os.path.dirname(_HERE), "tests", "golden", "poseidon_params.json")
    P = json.load(open(golden))
    m = lambda x: np.frombuffer(B.fr_to_mont_bytes(int(x, 16)), dtype=np.uint64)
    rc = np.concatenate([m(x) for row in P["round_constants"] for x in row])
    mds = np.concatenate([m(x) for row in P["mds"] for x in row])
    lib().oracle_poseidon_set_params(_p(np.ascontiguousarray(rc)), _p(np.ascontiguousarray(mds)))
    _poseidon_ready = True


def poseidon_hash(inputs_mont) -> np.ndarray:
    _poseidon_init()
    a = _u64(inputs_mont)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_poseidon_hash(_p(out), _p(a), ctypes.c_size_t(a.size // 4))
    return out


class MstC:
    """Merkle sum tree built by the C oracle (threaded): the whole tree as flat level-major arrays."""

    def __init__(self, names, balances: np.ndarray):
        """names: list of bytes; balances (n, n_cur) uint64."""
        _poseidon_init()
        n = len(names)
        bal = np.ascontiguousarray(balances, dtype=np.uint64)
        self.n_cur = bal.shape[1]
        self.depth = max(0, (n - 1).bit_length())
        offs = np.zeros(n + 1, dtype=np.uint32)
        offs[1:] = np.cumsum([len(x) for x in names])
        blob = np.frombuffer(b"".join(names), dtype=np.uint8) if offs[-1] else np.zeros(1, dtype=np.uint8)
        slots = 2 << self.depth
        self.hashes = np.zeros((slots, 4), dtype=np.uint64)
        self.balances = np.zeros((slots, self.n_cur, 4), dtype=np.uint64)
        self.unames = np.zeros((1 << self.depth, 4), dtype=np.uint64)
        lib().oracle_mst_build(_p(self.hashes), _p(self.balances), _p(self.unames), _p(np.ascontiguousarray(blob)), _p(offs), _p(bal), ctypes.c_size_t(n),
                               ctypes.c_uint32(self.n_cur), ctypes.c_uint32(self.depth))

    def level_offset(self, level: int) -> int:
        return sum(1 << (self.depth - l) for l in range(level))

    def node(self, level: int, index: int):
        o = self.level_offset(level) + index
        return self.hashes[o], self.balances[o]

    def root(self):
        return self.node(self.depth, 0)


def g1_fixed_base_mul(scalars) -> np.ndarray:
    """out[i] = scalars[i] * G (affine Montgomery, (n, 8) uint64): the data-parallel core of `ParamsKZG::setup`."""
    s = _u64(scalars)
    n = s.size // 4
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oracle_g1_fixed_base_mul(_p(out), _p(s), ctypes.c_size_t(n))
    return out
