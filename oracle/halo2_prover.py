"""CPU ORACLE (test infrastructure, NOT the product) -- halo2 `keygen_pk` data and `create_proof`
(KZG commitments, SHPLONK multi-open) restated for one circuit instance.

Reference call sites: zk_prover/src/circuits/utils.rs:75-76 (keygen), :94-102 (Blake2b transcript),
:171-178 (Keccak transcript).  The implementation lives in the un-vendored halo2_proofs 0.2.0 @
summa-dev/halo2#8386d6e; the algorithm below restates SURVEY.md A.4-A.10/A.13 (create_proof order and
RNG draw order, permutation / lookup / vanishing arguments, evaluate_h fold order, evaluation order,
SHPLONK rotation sets).  PINNING: proofs made here are accepted by the reference's own verifier
(contracts/src/InclusionVerifier.sol executed by oracle/yul.py) -- tests/test_oracle_prover.py.
Byte equality with the Rust prover is unpinned (it only ever runs with OsRng).

Vectors are numpy uint64 arrays of shape (len, 4): halo2curves' Montgomery memory layout; heavy
lifting (NTT, MSM, field vector ops) goes to oracle/halo2_cpu.c.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import bn254 as B
from . import cpu

R = B.R


# ------------------------------------------------------------------ small helpers
def m(x: int) -> np.ndarray:
    """Montgomery limbs (4,) of a python int."""
    return np.frombuffer(B.fr_to_mont_bytes(x % R), dtype=np.uint64).copy()


def um(limbs) -> int:
    return B.fr_from_mont_bytes(np.ascontiguousarray(limbs).tobytes())


def const_vec(x: int, n: int) -> np.ndarray:
    return np.tile(m(x), (n, 1))


def from_ints(xs: Sequence[int]) -> np.ndarray:
    return np.frombuffer(B.frs_to_bytes(xs), dtype=np.uint64).reshape(-1, 4).copy()


def to_ints(a: np.ndarray) -> List[int]:
    return B.frs_from_bytes(np.ascontiguousarray(a).tobytes())


def vmul(a, b):
    return cpu.par_fr_mul(a.reshape(-1), b.reshape(-1)).reshape(-1, 4)


def vadd(a, b):
    return cpu.par_fr_add(a.reshape(-1), b.reshape(-1)).reshape(-1, 4)


def vsub(a, b):
    return cpu.par_fr_sub(a.reshape(-1), b.reshape(-1)).reshape(-1, 4)


def vscale(a, s: int):
    return cpu.par_fr_scale(a.reshape(-1), m(s)).reshape(-1, 4)


def vadd_const(a, s: int):
    return cpu.par_fr_add_const(a.reshape(-1), m(s)).reshape(-1, 4)


def eval_poly(coeffs: np.ndarray, x: int) -> int:
    return um(cpu.par_fr_eval_poly(coeffs.reshape(-1), m(x)))


def rot(a: np.ndarray, r: int, scale: int = 1) -> np.ndarray:
    """values of p(omega^r X) on the same domain: index shift by r * scale."""
    return np.roll(a, -r * scale, axis=0) if r else a


def eval_expr(e, col, n: int) -> np.ndarray:
    """Evaluate an expression tree (oracle/sol_cs.py JSON) over whole columns; col(kind, c, rot) -> array."""
    k = e[0]
    if k == "const":
        return const_vec(int(e[1], 16), n)
    if k in ("advice", "fixed", "instance"):
        return col(k, e[1], e[2])
    if k == "neg":
        return vsub(np.zeros((n, 4), dtype=np.uint64), eval_expr(e[1], col, n))
    a, b = eval_expr(e[1], col, n), eval_expr(e[2], col, n)
    return vadd(a, b) if k == "add" else vsub(a, b) if k == "sub" else vmul(a, b)


class ProgramBuilder:
    """Flattens expression trees into the register program `oracle_expr_eval` runs row by row -- halo2's `GraphEvaluator`
    (SURVEY A.12): identical sub-expressions are computed once (`add_calculation` de-duplicates), every value lives in a
    per-thread `intermediates` slot.  Nodes: ("const", int) | ("col", index, rot) | ("neg", a) | ("add"|"sub"|"mul", a, b),
    where a / b are node ids returned by earlier calls."""
    OPS = {"const": 0, "col": 1, "add": 2, "sub": 3, "mul": 4, "neg": 5}

    def __init__(self):
        self.memo: Dict[tuple, int] = {}
        self.code: List[Tuple[int, int, int, int]] = []
        self.consts: List[int] = []
        self.const_ix: Dict[int, int] = {}

    def _emit(self, key, op, a, b):
        if key in self.memo:
            return self.memo[key]
        dst = len(self.code)
        self.code.append((op, dst, a, b))
        self.memo[key] = dst
        return dst

    def const(self, v: int) -> int:
        v %= R
        if v not in self.const_ix:
            self.const_ix[v] = len(self.consts)
            self.consts.append(v)
        return self._emit(("const", v), 0, self.const_ix[v], 0)

    def col(self, index: int, rot_: int = 0) -> int:
        return self._emit(("col", index, rot_), 1, index, rot_)

    def add(self, a, b): return self._emit(("add", min(a, b), max(a, b)), 2, a, b)
    def sub(self, a, b): return self._emit(("sub", a, b), 3, a, b)
    def mul(self, a, b): return self._emit(("mul", min(a, b), max(a, b)), 4, a, b)
    def neg(self, a): return self._emit(("neg", a), 5, a, 0)

    def tree(self, e, colmap) -> int:
        """expression tree of the constraint-system JSON; colmap(kind, column) -> column-table index"""
        k = e[0]
        if k == "const":
            return self.const(int(e[1], 16))
        if k in ("advice", "fixed", "instance"):
            return self.col(colmap(k, e[1]), e[2])
        if k == "neg":
            return self.neg(self.tree(e[1], colmap))
        a, b = self.tree(e[1], colmap), self.tree(e[2], colmap)
        return self.add(a, b) if k == "add" else self.sub(a, b) if k == "sub" else self.mul(a, b)

    def run(self, cols, out: int, n_rows: int, rot_scale: int) -> np.ndarray:
        code = np.array(self.code, dtype=np.int32).reshape(-1, 4)
        consts = from_ints(self.consts) if self.consts else np.zeros((0, 4), dtype=np.uint64)
        return cpu.expr_eval(cols, code, consts, len(self.code), out, n_rows, rot_scale)


class Params:
    """ParamsKZG (prover side) over the oracle MSM."""

    def __init__(self, k: int, g: np.ndarray, g_lagrange: np.ndarray, threads: int = 8):
        self.k, self.n, self.g, self.g_lagrange, self.threads = k, 1 << k, g, g_lagrange, threads

    @classmethod
    def read(cls, path: str, threads: int = 8) -> "Params":
        p = B.ParamsKZG.read(path)
        return cls(p.k, np.frombuffer(p.g_bytes, dtype=np.uint64).reshape(-1, 8), np.frombuffer(p.g_lagrange_bytes, dtype=np.uint64).reshape(-1, 8), threads)

    @classmethod
    def setup(cls, k: int, tau: int, threads: int = 8) -> "Params":
        """`ParamsKZG::setup(k, rng)` (utils.rs:70) with an explicit secret (UNSAFE test SRS): g[i] = [tau^i] G,
        g_lagrange[i] = [L_i(tau)] G with L_i(tau) = omega^i (tau^n - 1) / (n (tau - omega^i))."""
        n = 1 << k
        cpu.set_threads(threads)
        d = B.EvaluationDomain(3, k)
        g = cpu.g1_fixed_base_mul(cpu.fr_powers(m(tau), n).reshape(-1))
        w = cpu.fr_powers(m(d.omega), n)
        den = cpu.par_fr_batch_invert(vsub(const_vec(tau, n), w).reshape(-1)).reshape(n, 4)
        c = (pow(tau, n, R) - 1) * d.ifft_divisor % R
        lag = vscale(vmul(w, den), c)
        return cls(k, g, cpu.g1_fixed_base_mul(lag.reshape(-1)), threads)

    def _commit(self, bases, scalars) -> B.Point:
        out = cpu.best_multiexp(np.ascontiguousarray(scalars).reshape(-1), bases[: scalars.shape[0]].reshape(-1), self.threads)
        return B.g1_from_mont_bytes(out.tobytes())

    def commit(self, coeffs) -> B.Point:
        return self._commit(self.g, coeffs)

    def commit_lagrange(self, evals) -> B.Point:
        return self._commit(self.g_lagrange, evals)


# ------------------------------------------------------------------ keygen (the parts create_proof reads)
class ProvingKey:
    """What halo2's `ProvingKey` + `VerifyingKey` hold for `create_proof` (keygen_vk / keygen_pk)."""

    def __init__(self, params: Params, cs: dict, fixed: np.ndarray, sigma_mapping: Optional[List[List[Tuple[int, int]]]] = None,
                 transcript_repr: int = 0x10F28BC710A8BDD00DD701DF2F5FC4F5CCDB260238EBA6F819DB692F79DC3DC9, perm_cells=None):
        """fixed: (F, n, 4) Lagrange values of the fixed columns (after selector compression).
        sigma_mapping[j][i] = (column index in cs['permutation_columns'], row) that cell (j, i) is
        mapped to by the copy-constraint permutation (identity if None); perm_cells: the same in sparse form, rows
        (col, row, to_col, to_row) for the cells that moved (the format of tests/golden/mst_inclusion_assignment*.npz)."""
        self.cs = cs
        self.k, self.n = params.k, params.n
        n = self.n
        self.dom = cpu.Domain(cs["degree"], params.k, threads=params.threads)
        d = self.dom.d
        self.ext_n = 1 << d.extended_k
        self.rot_scale = 1 << (d.extended_k - params.k)
        self.transcript_repr = transcript_repr  # vk.hash_into: Blake2b of the pinned vk Debug string (an input)
        self.fixed_values = fixed
        self.fixed_polys = np.stack([self.dom.lagrange_to_coeff(f.reshape(-1)).reshape(n, 4) for f in fixed])
        self.fixed_cosets = np.stack([self.dom.coeff_to_extended(p.reshape(-1)).reshape(-1, 4) for p in self.fixed_polys])
        self.fixed_commitments = [params.commit_lagrange(f) for f in fixed]
        # permutation argument: sigma_j(omega^i) = delta^(col') * omega^(row')
        ncols = len(cs["permutation_columns"])
        self.omega_pows = cpu.fr_powers(m(d.omega), n)
        delta_pows = [pow(B.DELTA, j, R) for j in range(ncols)]
        sig = []
        for j in range(ncols):
            # identity permutation delta^j * omega^i, then the cells moved by copy constraints
            v = vscale(self.omega_pows, delta_pows[j])
            if perm_cells is not None:
                for (c, row, tc, trow) in perm_cells:
                    if c == j:
                        v[row] = m(delta_pows[tc] * pow(d.omega, int(trow), R) % R)
            elif sigma_mapping is not None:
                for i in range(n):
                    if sigma_mapping[j][i] != (j, i):
                        v[i] = m(delta_pows[sigma_mapping[j][i][0]] * pow(d.omega, sigma_mapping[j][i][1], R) % R)
            sig.append(v)
        self.sigma_values = np.stack(sig)
        self.sigma_polys = np.stack([self.dom.lagrange_to_coeff(s.reshape(-1)).reshape(n, 4) for s in sig])
        self.sigma_cosets = np.stack([self.dom.coeff_to_extended(p.reshape(-1)).reshape(-1, 4) for p in self.sigma_polys])
        self.sigma_commitments = [params.commit_lagrange(s) for s in sig]
        # l_0, l_last, l_blind -> extended; l_active_row = 1 - (l_last + l_blind)
        bf = cs["blinding_factors"]
        l0 = np.zeros((n, 4), dtype=np.uint64); l0[0] = m(1)
        l_blind = np.zeros((n, 4), dtype=np.uint64); l_blind[n - bf:] = m(1)
        l_last = np.zeros((n, 4), dtype=np.uint64); l_last[n - bf - 1] = m(1)
        ext = lambda v: self.dom.coeff_to_extended(self.dom.lagrange_to_coeff(v.reshape(-1))).reshape(-1, 4)
        self.l0, self.l_last = ext(l0), ext(l_last)
        l_blind_ext = ext(l_blind)
        self.l_active_row = vsub(const_vec(1, self.ext_n), vadd(self.l_last, l_blind_ext))
        # X on the extended coset: zeta * ext_omega^i
        self.x_coset = vscale(cpu.fr_powers(m(d.extended_omega), self.ext_n), d.g_coset)

    @classmethod
    def from_sparse(cls, params: Params, cs: dict, fixed_cells, fixed_cell_values, perm_cells, transcript_repr: int) -> "ProvingKey":
        """keygen output in the sparse form of tests/golden/mst_inclusion_assignment*.npz, expanded to k = params.k"""
        fixed = np.zeros((cs["num_fixed_columns"], params.n, 4), dtype=np.uint64)
        fc = np.asarray(fixed_cells)
        fixed[fc[:, 0], fc[:, 1]] = np.asarray(fixed_cell_values, dtype=np.uint64)
        return cls(params, cs, fixed, None, transcript_repr, perm_cells=[tuple(int(x) for x in r) for r in np.asarray(perm_cells)])


# ------------------------------------------------------------------ lookup: permute_expression_pair
def permute_expression_pair(inp: List[int], tab: List[int], usable_rows: int) -> Tuple[List[int], List[int]]:
    permuted_input = sorted(inp[:usable_rows])
    leftover: Dict[int, int] = {}
    for v in tab[:usable_rows]:
        leftover[v] = leftover.get(v, 0) + 1
    permuted_table = [0] * usable_rows
    repeated = []
    for row, v in enumerate(permuted_input):
        if row == 0 or v != permuted_input[row - 1]:
            permuted_table[row] = v
            if leftover.get(v, 0) == 0:
                raise ValueError("lookup input value not in table (ConstraintSystemFailure)")
            leftover[v] -= 1
        else:
            repeated.append(row)
    for v in sorted(leftover):
        for _ in range(leftover[v]):
            permuted_table[repeated.pop()] = v
    assert not repeated
    return permuted_input, permuted_table


# ------------------------------------------------------------------ SHPLONK
def lagrange_interpolate(points: List[int], evals: List[int]) -> List[int]:
    """coefficients (low -> high) of the polynomial of degree < len(points) through (points, evals)."""
    n = len(points)
    out = [0] * n
    for j in range(n):
        num = [1]
        den = 1
        for kk in range(n):
            if kk == j:
                continue
            num = [(a - points[kk] * b) % R for a, b in zip([0] + num, num + [0])]
            den = den * (points[j] - points[kk]) % R
        s = evals[j] * pow(den, -1, R) % R
        for i in range(len(num)):
            out[i] = (out[i] + num[i] * s) % R
    return out


def poly_eval_small(c: List[int], x: int) -> int:
    acc = 0
    for a in reversed(c):
        acc = (acc * x + a) % R
    return acc


def shplonk_create_proof(params: Params, transcript, queries: List[Tuple[int, int, np.ndarray, int]]):
    """queries: list of (poly_id, point, poly coefficients (n,4), eval) in halo2's query order."""
    n = params.n
    y = transcript.squeeze_challenge()
    # construct_intermediate_sets: polynomials in first-appearance order, rotation sets in first-appearance order,
    # points inside a set in ascending field order (BTreeSet<Fr>)
    poly_points: Dict[int, set] = {}
    poly_order: List[int] = []
    polys: Dict[int, np.ndarray] = {}
    evals: Dict[Tuple[int, int], int] = {}
    super_points = set()
    for pid, pt, poly, ev in queries:
        super_points.add(pt)
        if pid not in poly_points:
            poly_points[pid] = set()
            poly_order.append(pid)
            polys[pid] = poly
        poly_points[pid].add(pt)
        evals[(pid, pt)] = ev
    sets: List[Tuple[Tuple[int, ...], List[int]]] = []
    for pid in poly_order:
        key = tuple(sorted(poly_points[pid]))
        for s_key, members in sets:
            if s_key == key:
                members.append(pid)
                break
        else:
            sets.append((key, [pid]))
    v = transcript.squeeze_challenge()

    def div_by_vanishing(poly: np.ndarray, roots: Sequence[int]) -> np.ndarray:
        cur = poly.reshape(-1)
        for rt in roots:
            cur = cpu.fr_kate_division(cur, m(rt))
        out = np.zeros((n, 4), dtype=np.uint64)
        out[: cur.size // 4] = cur.reshape(-1, 4)
        return out

    # quotient contributions
    h_x = np.zeros((n, 4), dtype=np.uint64)
    v_pow = 1
    low_deg: Dict[Tuple[int, int], List[int]] = {}
    for si, (pts, members) in enumerate(sets):
        n_x = np.zeros((n, 4), dtype=np.uint64)
        y_pow = 1
        for pid in members:
            r_coeffs = lagrange_interpolate(list(pts), [evals[(pid, p)] for p in pts])
            low_deg[(si, pid)] = r_coeffs
            num = polys[pid].copy()
            num[: len(r_coeffs)] = vsub(num[: len(r_coeffs)], from_ints(r_coeffs))
            n_x = vadd(n_x, vscale(num, y_pow))
            y_pow = y_pow * y % R
        q_i = div_by_vanishing(n_x, pts)
        h_x = vadd(h_x, vscale(q_i, v_pow))
        v_pow = v_pow * v % R
    transcript.write_point(params.commit(h_x))
    u = transcript.squeeze_challenge()
    sorted_super = sorted(super_points)
    l_x = np.zeros((n, 4), dtype=np.uint64)
    v_pow = 1
    z_diffs = []
    for si, (pts, members) in enumerate(sets):
        z_i = 1
        for p in sorted_super:
            if p not in pts:
                z_i = z_i * (u - p) % R
        z_diffs.append(z_i)
        inner = np.zeros((n, 4), dtype=np.uint64)
        y_pow = 1
        for pid in members:
            r_eval = poly_eval_small(low_deg[(si, pid)], u)
            p_x = polys[pid].copy()
            p_x[0] = m((um(p_x[0]) - r_eval) % R)
            inner = vadd(inner, vscale(p_x, y_pow))
            y_pow = y_pow * y % R
        l_x = vadd(l_x, vscale(inner, z_i * v_pow % R))
        v_pow = v_pow * v % R
    zt_eval = 1
    for p in sorted_super:
        zt_eval = zt_eval * (u - p) % R
    l_x = vsub(l_x, vscale(h_x, zt_eval))
    assert eval_poly(l_x, u) == 0, "SHPLONK linearisation does not vanish at u"
    hq = div_by_vanishing(l_x, [u])
    hq = vscale(hq, pow(z_diffs[0], -1, R))
    transcript.write_point(params.commit(hq))


# ------------------------------------------------------------------ evaluate_h
def _h_numerator_columnwise(cs, pk, ext_col, perm_sets, lk_cosets, theta, beta, gamma, y, ext_n, rs, bf):
    """The quotient numerator term by term over whole columns: the readable form (one pass over the extended domain per operator)."""
    h = np.zeros((ext_n, 4), dtype=np.uint64)
    fold = lambda acc, term: vadd(vscale(acc, y), term)
    for g in cs["gates"]:
        h = fold(h, eval_expr(g, ext_col, ext_n))
    one = const_vec(1, ext_n)
    if perm_sets:
        first, last = perm_sets[0], perm_sets[-1]
        h = fold(h, vmul(vsub(one, first["coset"]), pk.l0))
        h = fold(h, vmul(vsub(vmul(last["coset"], last["coset"]), last["coset"]), pk.l_last))
        for i in range(1, len(perm_sets)):
            h = fold(h, vmul(vsub(perm_sets[i]["coset"], rot(perm_sets[i - 1]["coset"], -(bf + 1), rs)), pk.l0))
        beta_x = vscale(pk.x_coset, beta)   # beta * X on the extended coset: X = zeta * ext_omega^idx
        cur_delta = 1
        for s in perm_sets:
            left = rot(s["coset"], 1, rs)
            right = s["coset"]
            for j, kc in enumerate(s["cols"]):
                vals = ext_col(kc[0], kc[1], 0)
                left = vmul(left, vadd_const(vadd(vals, vscale(pk.sigma_cosets[s["first"] + j], beta)), gamma))
                right = vmul(right, vadd_const(vadd(vals, vscale(beta_x, cur_delta)), gamma))
                cur_delta = cur_delta * B.DELTA % R
            h = fold(h, vmul(vsub(left, right), pk.l_active_row))
    for (zc, ic, tc), lkdef in zip(lk_cosets, cs["lookups"]):

        def compress_ext(exprs):
            acc = np.zeros((ext_n, 4), dtype=np.uint64)
            for e in exprs:
                acc = vadd(vscale(acc, theta), eval_expr(e, ext_col, ext_n))
            return acc
        tvi = vmul(vadd_const(compress_ext(lkdef["input"]), beta), vadd_const(compress_ext(lkdef["table"]), gamma))
        a_minus_s = vsub(ic, tc)
        h = fold(h, vmul(vsub(one, zc), pk.l0))
        h = fold(h, vmul(vsub(vmul(zc, zc), zc), pk.l_last))
        h = fold(h, vmul(vsub(vmul(rot(zc, 1, rs), vmul(vadd_const(ic, beta), vadd_const(tc, gamma))), vmul(zc, tvi)), pk.l_active_row))
        h = fold(h, vmul(a_minus_s, pk.l0))
        h = fold(h, vmul(vmul(a_minus_s, vsub(ic, rot(ic, -1, rs))), pk.l_active_row))
    return h


def _h_numerator_program(cs, pk, advice_cosets, inst_coset, perm_sets, lk_cosets, theta, beta, gamma, y, ext_n, rs, bf):
    """The same numerator the way halo2's `Evaluator::evaluate_h` computes it: ONE pass over the extended domain, rows split across
    threads, every row running the flattened calculation list (`GraphEvaluator`); terms folded by Horner in y in the same order."""
    A, F, P = cs["num_advice_columns"], cs["num_fixed_columns"], len(cs["permutation_columns"])
    n_sets = len(perm_sets)
    cols = list(advice_cosets) + [pk.fixed_cosets[c] for c in range(F)] + [inst_coset] + [pk.sigma_cosets[j] for j in range(P)]
    E_SIGMA = A + F + 1
    E_PZ = len(cols); cols += [s_["coset"] for s_ in perm_sets]
    E_L0 = len(cols); cols += [pk.l0, pk.l_last, pk.l_active_row, pk.x_coset]
    E_LLAST, E_LACT, E_X = E_L0 + 1, E_L0 + 2, E_L0 + 3
    E_LK = len(cols)
    for trio in lk_cosets:
        cols += list(trio)
    colmap = lambda kind, c: c if kind == "advice" else A + c if kind == "fixed" else A + F
    pb = ProgramBuilder()
    terms = [pb.tree(g, colmap) for g in cs["gates"]]
    one, l0, llast, lact = pb.const(1), pb.col(E_L0), pb.col(E_LLAST), pb.col(E_LACT)
    cb, cg = pb.const(beta), pb.const(gamma)
    if n_sets:
        Z = lambda s_, r: pb.col(E_PZ + s_, r)
        terms.append(pb.mul(pb.sub(one, Z(0, 0)), l0))
        zl = Z(n_sets - 1, 0)
        terms.append(pb.mul(pb.sub(pb.mul(zl, zl), zl), llast))
        for i in range(1, n_sets):
            terms.append(pb.mul(pb.sub(Z(i, 0), Z(i - 1, -(bf + 1))), l0))
        cur_delta = 1
        for si, s_ in enumerate(perm_sets):
            left, right = Z(si, 1), Z(si, 0)
            for j, kc in enumerate(s_["cols"]):
                val = pb.col(colmap(kc[0], kc[1]))
                left = pb.mul(left, pb.add(pb.add(val, pb.mul(cb, pb.col(E_SIGMA + s_["first"] + j))), cg))
                right = pb.mul(right, pb.add(pb.add(val, pb.mul(pb.const(cur_delta * beta % R), pb.col(E_X))), cg))
                cur_delta = cur_delta * B.DELTA % R
            terms.append(pb.mul(pb.sub(left, right), lact))
    cth = pb.const(theta)
    for li, lkdef in enumerate(cs["lookups"]):
        z0, z1 = pb.col(E_LK + 3 * li), pb.col(E_LK + 3 * li, 1)
        a0, am1, s0 = pb.col(E_LK + 3 * li + 1), pb.col(E_LK + 3 * li + 1, -1), pb.col(E_LK + 3 * li + 2)

        def compress(exprs):
            acc = None
            for e in exprs:
                t = pb.tree(e, colmap)
                acc = t if acc is None else pb.add(pb.mul(acc, cth), t)
            return acc
        cin, ctab = compress(lkdef["input"]), compress(lkdef["table"])
        a_minus_s = pb.sub(a0, s0)
        terms.append(pb.mul(pb.sub(one, z0), l0))
        terms.append(pb.mul(pb.sub(pb.mul(z0, z0), z0), llast))
        terms.append(pb.mul(pb.sub(pb.mul(z1, pb.mul(pb.add(a0, cb), pb.add(s0, cg))), pb.mul(z0, pb.mul(pb.add(cin, cb), pb.add(ctab, cg)))), lact))
        terms.append(pb.mul(a_minus_s, l0))
        terms.append(pb.mul(pb.mul(a_minus_s, pb.sub(a0, am1)), lact))
    cy = pb.const(y)
    acc = None
    for t in terms:
        acc = t if acc is None else pb.add(pb.mul(acc, cy), t)
    return pb.run(cols, acc, ext_n, rs)


# ------------------------------------------------------------------ create_proof
def create_proof(params: Params, pk: ProvingKey, instances: List[int], advice: np.ndarray, rng, transcript, trace: Optional[dict] = None):
    """advice: (A, n, 4) assigned advice columns (rows >= n - 6 are overwritten with blinding).
    rng must offer next_fr() / fill_bytes() (oracle.chacha.ChaCha20Rng).  Returns nothing; the proof is
    transcript.finalize().  `trace`, if given, receives intermediate objects for parity tests."""
    cs, dom, n = pk.cs, pk.dom, pk.n
    d = dom.d
    bf = cs["blinding_factors"]
    usable = n - (bf + 1)
    ext_n, rs = pk.ext_n, pk.rot_scale
    A = cs["num_advice_columns"]
    tr = trace if trace is not None else {}

    transcript.common_scalar(pk.transcript_repr)
    for v in instances:
        transcript.common_scalar(v)
    inst_vals = np.zeros((n, 4), dtype=np.uint64)
    inst_vals[: len(instances)] = from_ints(instances)
    assert len(instances) <= usable
    inst_poly = dom.lagrange_to_coeff(inst_vals.reshape(-1)).reshape(n, 4)

    # advice: blinding rows, blinds (drawn, unused by KZG), commitments
    advice = advice.copy()
    for c in range(A):
        advice[c, usable:] = from_ints([rng.next_fr() for _ in range(n - usable)])
    _advice_blinds = [rng.next_fr() for _ in range(A)]
    advice_polys = np.stack([dom.lagrange_to_coeff(a.reshape(-1)).reshape(n, 4) for a in advice])
    advice_comms = [params.commit_lagrange(a) for a in advice]
    for p in advice_comms:
        transcript.write_point(p)
    tr["advice_comms"] = advice_comms
    theta = transcript.squeeze_challenge()

    def lagrange_col(kind, c, r):
        src = advice[c] if kind == "advice" else pk.fixed_values[c] if kind == "fixed" else inst_vals
        return rot(src, r)

    # lookups: compress, permute, commit
    lookups = []
    for lk in cs["lookups"]:
        def compress(exprs):
            acc = np.zeros((n, 4), dtype=np.uint64)
            for e in exprs:
                acc = vadd(vscale(acc, theta), eval_expr(e, lagrange_col, n))
            return acc
        c_in, c_tab = compress(lk["input"]), compress(lk["table"])
        if n <= 4096:   # the readable python twin (the C form is checked against it in tests/test_oracle_circuit.py)
            p_in, p_tab = permute_expression_pair(to_ints(c_in), to_ints(c_tab), usable)
            p_in_u, p_tab_u = from_ints(p_in), from_ints(p_tab)
        else:
            p_in_u, p_tab_u = cpu.permute_expression_pair(c_in, c_tab, usable)
        p_in_v = np.concatenate([p_in_u, from_ints([rng.next_fr() for _ in range(bf + 1)])])
        p_tab_v = np.concatenate([p_tab_u, from_ints([rng.next_fr() for _ in range(bf + 1)])])
        in_poly = dom.lagrange_to_coeff(p_in_v.reshape(-1)).reshape(n, 4)
        _b = rng.next_fr()
        in_comm = params.commit_lagrange(p_in_v)
        tab_poly = dom.lagrange_to_coeff(p_tab_v.reshape(-1)).reshape(n, 4)
        _b = rng.next_fr()
        tab_comm = params.commit_lagrange(p_tab_v)
        transcript.write_point(in_comm)
        transcript.write_point(tab_comm)
        lookups.append(dict(c_in=c_in, c_tab=c_tab, p_in=p_in_v, p_tab=p_tab_v, in_poly=in_poly, tab_poly=tab_poly))
    beta = transcript.squeeze_challenge()
    gamma = transcript.squeeze_challenge()

    # permutation argument
    chunk = cs["degree"] - 2
    pcols = cs["permutation_columns"]
    col_values = lambda kc: advice[kc[1]] if kc[0] == "advice" else pk.fixed_values[kc[1]] if kc[0] == "fixed" else inst_vals
    omega_vec = pk.omega_pows
    deltaomega = 1  # delta^col
    last_z = 1
    perm_sets = []
    for s0 in range(0, len(pcols), chunk):
        cols = pcols[s0:s0 + chunk]
        mod = const_vec(1, n)
        for j, kc in enumerate(cols):
            mod = vmul(mod, vadd_const(vadd(vscale(pk.sigma_values[s0 + j], beta), col_values(kc)), gamma))
        mod = cpu.par_fr_batch_invert(mod.reshape(-1)).reshape(n, 4)
        for kc in cols:
            mod = vmul(mod, vadd_const(vadd(vscale(omega_vec, deltaomega * beta % R), col_values(kc)), gamma))
            deltaomega = deltaomega * B.DELTA % R
        z = cpu.fr_running_product(mod.reshape(-1), m(last_z), n).reshape(n, 4)
        z[n - bf:] = from_ints([rng.next_fr() for _ in range(bf)])
        last_z = um(z[n - (bf + 1)])
        _b = rng.next_fr()
        comm = params.commit_lagrange(z)
        z_poly = dom.lagrange_to_coeff(z.reshape(-1)).reshape(n, 4)
        z_coset = dom.coeff_to_extended(z_poly.reshape(-1)).reshape(-1, 4)
        transcript.write_point(comm)
        perm_sets.append(dict(poly=z_poly, coset=z_coset, cols=cols, first=s0))
    tr["perm_z"] = [s["poly"] for s in perm_sets]

    # lookup products
    for lk in lookups:
        den = vmul(vadd_const(lk["p_in"], beta), vadd_const(lk["p_tab"], gamma))
        den = cpu.par_fr_batch_invert(den.reshape(-1)).reshape(n, 4)
        prod = vmul(den, vmul(vadd_const(lk["c_in"], beta), vadd_const(lk["c_tab"], gamma)))
        z = cpu.fr_running_product(prod.reshape(-1), m(1), n).reshape(n, 4)
        z[n - bf:] = from_ints([rng.next_fr() for _ in range(bf)])
        _b = rng.next_fr()
        comm = params.commit_lagrange(z)
        lk["z_poly"] = dom.lagrange_to_coeff(z.reshape(-1)).reshape(n, 4)
        transcript.write_point(comm)

    # vanishing argument: random polynomial (one ChaCha20 child stream seeded from rng = the 1-thread case of the fork)
    from .chacha import ChaCha20Rng
    child = ChaCha20Rng(rng.fill_bytes(32))
    random_poly = cpu.chacha_fr_fill(child.key, 0, n)   # coefficient i = `Fr::random` draw i = keystream block i
    _b = rng.next_fr()
    transcript.write_point(params.commit(random_poly))
    y = transcript.squeeze_challenge()

    # ---- evaluate_h on the extended coset ----
    advice_cosets = [dom.coeff_to_extended(p.reshape(-1)).reshape(-1, 4) for p in advice_polys]
    inst_coset = dom.coeff_to_extended(inst_poly.reshape(-1)).reshape(-1, 4)

    def ext_col(kind, c, r):
        src = advice_cosets[c] if kind == "advice" else pk.fixed_cosets[c] if kind == "fixed" else inst_coset
        return rot(src, r, rs)

    lk_cosets = [[dom.coeff_to_extended(lk[nm].reshape(-1)).reshape(-1, 4) for nm in ("z_poly", "in_poly", "tab_poly")] for lk in lookups]
    if tr.get("columnwise_h"):
        h = _h_numerator_columnwise(cs, pk, ext_col, perm_sets, lk_cosets, theta, beta, gamma, y, ext_n, rs, bf)
    else:
        h = _h_numerator_program(cs, pk, advice_cosets, inst_coset, perm_sets, lk_cosets, theta, beta, gamma, y, ext_n, rs, bf)
    tr["h_numerator_ext"] = h

    # ---- quotient: divide by t(X), back to coefficients, 5 pieces ----
    h = dom.divide_by_vanishing_poly(h.reshape(-1))
    h_coeff = dom.extended_to_coeff(h).reshape(-1, 4)
    pieces = [h_coeff[i * n:(i + 1) * n] for i in range(d.quotient_poly_degree)]
    _hb = [rng.next_fr() for _ in pieces]
    for p in pieces:
        transcript.write_point(params.commit(p))
    tr["h_pieces"] = pieces
    x = transcript.squeeze_challenge()
    xn = pow(x, n, R)
    rotx = lambda r: x * pow(d.omega, r, R) % R

    # ---- evaluations (SURVEY A.5 order) ----
    advice_evals = [eval_poly(advice_polys[c], rotx(r)) for c, r in cs["advice_queries"]]
    for e in advice_evals:
        transcript.write_scalar(e)
    fixed_evals = [eval_poly(pk.fixed_polys[c], rotx(r)) for c, r in cs["fixed_queries"]]
    for e in fixed_evals:
        transcript.write_scalar(e)
    h_folded = np.zeros((n, 4), dtype=np.uint64)
    for p in reversed(pieces):
        h_folded = vadd(vscale(h_folded, xn), p)
    random_eval = eval_poly(random_poly, x)
    transcript.write_scalar(random_eval)
    sigma_evals = [eval_poly(p, x) for p in pk.sigma_polys]
    for e in sigma_evals:
        transcript.write_scalar(e)
    x_next, x_prev, x_last = rotx(1), rotx(-1), rotx(-(bf + 1))
    perm_evals = []
    for i, s in enumerate(perm_sets):
        ev = [eval_poly(s["poly"], x), eval_poly(s["poly"], x_next)]
        if i != len(perm_sets) - 1:
            ev.append(eval_poly(s["poly"], x_last))
        for e in ev:
            transcript.write_scalar(e)
        perm_evals.append(ev)
    lk_evals = []
    for lk in lookups:
        ev = [eval_poly(lk["z_poly"], x), eval_poly(lk["z_poly"], x_next), eval_poly(lk["in_poly"], x),
              eval_poly(lk["in_poly"], x_prev), eval_poly(lk["tab_poly"], x)]
        for e in ev:
            transcript.write_scalar(e)
        lk_evals.append(ev)

    # ---- multi-open queries in halo2's order ----
    q: List[Tuple[int, int, np.ndarray, int]] = []
    pid = {}

    def ident(tag):
        return pid.setdefault(tag, len(pid))
    for (c, r), e in zip(cs["advice_queries"], advice_evals):
        q.append((ident(("advice", c)), rotx(r), advice_polys[c], e))
    for i, s in enumerate(perm_sets):
        q.append((ident(("permz", i)), x, s["poly"], perm_evals[i][0]))
        q.append((ident(("permz", i)), x_next, s["poly"], perm_evals[i][1]))
    for i in reversed(range(len(perm_sets) - 1)):
        q.append((ident(("permz", i)), x_last, perm_sets[i]["poly"], perm_evals[i][2]))
    for i, lk in enumerate(lookups):
        ev = lk_evals[i]
        q.append((ident(("lkz", i)), x, lk["z_poly"], ev[0]))
        q.append((ident(("lkin", i)), x, lk["in_poly"], ev[2]))
        q.append((ident(("lktab", i)), x, lk["tab_poly"], ev[4]))
        q.append((ident(("lkin", i)), x_prev, lk["in_poly"], ev[3]))
        q.append((ident(("lkz", i)), x_next, lk["z_poly"], ev[1]))
    for (c, r), e in zip(cs["fixed_queries"], fixed_evals):
        q.append((ident(("fixed", c)), rotx(r), pk.fixed_polys[c], e))
    for j, e in enumerate(sigma_evals):
        q.append((ident(("sigma", j)), x, pk.sigma_polys[j], e))
    q.append((ident("h"), x, h_folded, eval_poly(h_folded, x)))
    q.append((ident("random"), x, random_poly, random_eval))
    shplonk_create_proof(params, transcript, q)
