"""CPU ORACLE (test infrastructure) -- `MstInclusionCircuit<LEVELS, N_CURRENCIES, N_BYTES>` synthesised
the way halo2's SimpleFloorPlanner + keygen do it, so the oracle owns the *real* fixed columns, copy
constraints and witness of the reference circuit.

Follows zk_prover/src/circuits/merkle_sum_tree.rs:141-207 (configure), :228-520 (synthesize),
circuits/traits.rs:9-52, chips/merkle_sum_tree.rs:107-227, chips/range/range_check.rs:93-153,
chips/poseidon/hash.rs:75-87 and, for the un-vendored halo2 pieces (SURVEY A.11 + crate knowledge):
halo2_gadgets Pow5Chip region layout (initial state / add input / permute state), SimpleFloorPlanner
region placement (a region starts at the max next-free row of the columns it uses), per-region constant
assignment into the first constants column, selector compression (`compress_selectors::process`) and
the permutation `Assembly::copy` cycle merge.
PINNED by the reference's verifying key: all 11 fixed_comms and 6 permutation_comms of
contracts/src/InclusionVerifier.sol:238-271 are reproduced bit-exactly (tests/test_oracle_circuit.py).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Tuple

from . import bn254 as B
from . import mst as M

R = B.R
# column ids
A0, A1, A2 = ("advice", 0), ("advice", 1), ("advice", 2)
F = [("fixed", i) for i in range(5)]
INST = ("instance", 0)
# selectors in creation order (merkle_sum_tree.rs:149-152, Pow5Chip::configure x2)
S_BOOL_SWAP, S_SUM, S_LOOKUP, S_FULL_E, S_PARTIAL_E, S_PAD_E, S_FULL_M, S_PARTIAL_M, S_PAD_M = range(9)
SELECTOR_MAX_DEGREE = {S_BOOL_SWAP: 3, S_SUM: 2, S_LOOKUP: 0, S_FULL_E: 6, S_PARTIAL_E: 6, S_PAD_E: 2, S_FULL_M: 6, S_PARTIAL_M: 6, S_PAD_M: 2}
PERM_COLUMNS = [F[2], A0, A1, F[3], A2, INST]  # order of enable_equality calls


class Cell:
    __slots__ = ("col", "row", "value")

    def __init__(self, col, row, value):
        self.col, self.row, self.value = col, row, value


class Layouter:
    """SingleChipLayouter: regions are measured first, then placed at the max next-free row of their columns."""

    def __init__(self, n: int):
        self.n = n
        self.next_free: Dict[object, int] = {}
        self.advice = [dict() for _ in range(3)]
        self.fixed = [dict() for _ in range(5)]
        self.selectors = [set() for _ in range(9)]
        self.copies: List[Tuple[object, int, object, int]] = []
        self.region_starts: List[int] = []

    def region(self, fn):
        shape = _Region(self, None)
        fn(shape)
        start = 0
        for c in shape.columns:
            start = max(start, self.next_free.get(c, 0))
        for c in shape.columns:
            self.next_free[c] = start + shape.rows
        self.region_starts.append(start)
        reg = _Region(self, start)
        result = fn(reg)
        # SingleChipLayouter::assign_region: this region's constants go, in order, into the first
        # constants column (fixed 2) at that column's next free row, one row each
        for value, cell in reg.constants:
            row = self.next_free.get(F[2], 0)
            self.fixed[2][row] = value
            self.copies.append((F[2], row, cell.col, cell.row))
            self.next_free[F[2]] = row + 1
        return result

    def constrain_instance(self, cell: Cell, row: int):
        self.copies.append((cell.col, cell.row, INST, row))

    def finish(self):
        pass


class _Region:
    def __init__(self, lay: Layouter, start: Optional[int]):
        self.lay, self.start = lay, start
        self.columns, self.rows = [], 0
        self.constants: List[Tuple[int, Cell]] = []

    def _touch(self, col, off):
        if col not in self.columns:
            self.columns.append(col)
        self.rows = max(self.rows, off + 1)

    def enable(self, sel: int, off: int):
        self._touch(("selector", sel), off)
        if self.start is not None:
            self.lay.selectors[sel].add(self.start + off)

    def assign_advice(self, col, off: int, value: int) -> Cell:
        self._touch(col, off)
        if self.start is None:
            return Cell(col, off, value)
        self.lay.advice[col[1]][self.start + off] = value % R
        return Cell(col, self.start + off, value % R)

    def assign_fixed(self, col, off: int, value: int):
        self._touch(col, off)
        if self.start is not None:
            self.lay.fixed[col[1]][self.start + off] = value % R

    def copy_advice(self, src: Cell, col, off: int) -> Cell:
        c = self.assign_advice(col, off, src.value)
        if self.start is not None:
            self.lay.copies.append((c.col, c.row, src.col, src.row))
        return c

    def assign_advice_from_constant(self, col, off: int, value: int) -> Cell:
        c = self.assign_advice(col, off, value)
        if self.start is not None:
            self.constants.append((value % R, c))
        return c

    def constrain_constant(self, cell: Cell, value: int):
        if self.start is not None:
            self.constants.append((value % R, cell))


# ------------------------------------------------------------------ chips
def assign_value(lay: Layouter, value: int, col) -> Cell:
    return lay.region(lambda r: r.assign_advice(col, 0, value))


def poseidon_hash_chip(lay: Layouter, inputs: List[Cell], s_full: int, s_partial: int, s_pad: int) -> Cell:
    """halo2_gadgets Hash<_, _, S, ConstantLength<L>, 2, 1>::init(..).hash(..) with Pow5Chip."""
    L = len(inputs)
    state = lay.region(lambda r: [r.assign_advice_from_constant(A0, 0, 0), r.assign_advice_from_constant(A1, 0, (L << 64) % R)])

    def add_input(r: _Region, st, word):
        r.enable(s_pad, 1)
        init = [r.copy_advice(st[0], A0, 0), r.copy_advice(st[1], A1, 0)]
        inp = r.copy_advice(word, A0, 1)
        return [r.assign_advice(A0, 2, init[0].value + inp.value), r.assign_advice(A1, 2, init[1].value)]

    def permute(r: _Region, st):
        cur = [r.copy_advice(st[0], A0, 0), r.copy_advice(st[1], A1, 0)]
        half_f, half_p = M.R_F // 2, M.R_P // 2

        def full(cur, rnd, off):
            r.enable(s_full, off)
            r.assign_fixed(F[0], off, M.RC[rnd][0])
            r.assign_fixed(F[1], off, M.RC[rnd][1])
            nxt = M.mds_mul([pow((cur[0].value + M.RC[rnd][0]) % R, 5, R), pow((cur[1].value + M.RC[rnd][1]) % R, 5, R)])
            return [r.assign_advice(A0, off + 1, nxt[0]), r.assign_advice(A1, off + 1, nxt[1])]

        def partial(cur, rnd, off):
            r.enable(s_partial, off)
            r.assign_fixed(F[0], off, M.RC[rnd][0])
            r.assign_fixed(F[1], off, M.RC[rnd][1])
            r0 = pow((cur[0].value + M.RC[rnd][0]) % R, 5, R)
            r1 = (cur[1].value + M.RC[rnd][1]) % R
            r.assign_advice(A2, off, r0)
            mid = M.mds_mul([r0, r1])
            r.assign_fixed(F[2], off, M.RC[rnd + 1][0])
            r.assign_fixed(F[3], off, M.RC[rnd + 1][1])
            nxt = M.mds_mul([pow((mid[0] + M.RC[rnd + 1][0]) % R, 5, R), (mid[1] + M.RC[rnd + 1][1]) % R])
            return [r.assign_advice(A0, off + 1, nxt[0]), r.assign_advice(A1, off + 1, nxt[1])]

        for i in range(half_f):
            cur = full(cur, i, i)
        for i in range(half_p):
            cur = partial(cur, half_f + 2 * i, half_f + i)
        for i in range(half_f):
            cur = full(cur, half_f + 2 * half_p + i, half_f + half_p + i)
        return cur

    for word in inputs:
        state = lay.region(lambda r, st=state, w=word: add_input(r, st, w))
        state = lay.region(lambda r, st=state: permute(r, st))
    return state[0]


def range_check(lay: Layouter, value: Cell, n_bytes: int):
    def body(r: _Region):
        for i in range(n_bytes):
            r.enable(S_LOOKUP, i)
        z = r.copy_advice(value, A0, 0)
        inv256 = pow(256, -1, R)
        v = z.value
        # decompose_fp_to_bytes: little-endian bytes of the canonical value, first n_bytes
        bs = list(int(v).to_bytes(32, "little")[:n_bytes])
        cur = z
        for i, b in enumerate(bs):
            cur = r.assign_advice(A0, i + 1, (cur.value - b) * inv256 % R)
        r.constrain_constant(cur, 0)
    lay.region(body)


def swap_hashes(lay: Layouter, current: Cell, sibling: Cell, bit: Cell):
    def body(r: _Region):
        r.enable(S_BOOL_SWAP, 0)
        l1 = r.copy_advice(current, A0, 0)
        r1 = r.copy_advice(sibling, A1, 0)
        sb = r.copy_advice(bit, A2, 0)
        lv, rv = (l1.value, r1.value) if sb.value == 0 else (r1.value, l1.value)
        return r.assign_advice(A0, 1, lv), r.assign_advice(A1, 1, rv)
    return lay.region(body)


def sum_balances(lay: Layouter, cur: Cell, elem: Cell) -> Cell:
    def body(r: _Region):
        r.enable(S_SUM, 0)
        a = r.copy_advice(cur, A0, 0)
        b = r.copy_advice(elem, A1, 0)
        return r.assign_advice(A2, 0, a.value + b.value)
    return lay.region(body)


def synthesize(k: int, proof: dict, levels: int, n_currencies: int, n_bytes: int = 8) -> Layouter:
    lay = Layouter(1 << k)
    pe = (S_FULL_E, S_PARTIAL_E, S_PAD_E)
    pm = (S_FULL_M, S_PARTIAL_M, S_PAD_M)
    pre = proof["entry"].preimage()
    username = assign_value(lay, pre[0], A0)
    balances = [assign_value(lay, pre[1 + i], A1) for i in range(n_currencies)]
    current_hash = poseidon_hash_chip(lay, [username] + balances, *pe)
    lay.constrain_instance(current_hash, 0)

    def table(r: _Region):
        for i in range(256):
            r.assign_fixed(F[4], i, i)
    lay.region(table)
    for level in range(levels):
        sib_bal: List[Cell] = []
        if level == 0:
            sp = proof["sibling_leaf_node_hash_preimage"]
            su = assign_value(lay, sp[0], A0)
            sib_bal = [assign_value(lay, sp[1 + c], A1) for c in range(n_currencies)]
            sibling_hash = poseidon_hash_chip(lay, [su] + sib_bal, *pe)
            for c in range(n_currencies):
                range_check(lay, balances[c], n_bytes)
                range_check(lay, sib_bal[c], n_bytes)
        else:
            sp = proof["sibling_middle_node_hash_preimages"][level - 1]
            sib_bal = [assign_value(lay, sp[c], A1) for c in range(n_currencies)]
            lh = assign_value(lay, sp[n_currencies], A2)
            rh = assign_value(lay, sp[n_currencies + 1], A2)
            sibling_hash = poseidon_hash_chip(lay, sib_bal + [lh, rh], *pm)
            for c in range(n_currencies):
                range_check(lay, sib_bal[c], n_bytes)
        bit = assign_value(lay, proof["path_indices"][level], A0)
        left, right = swap_hashes(lay, current_hash, sibling_hash, bit)
        nxt = [sum_balances(lay, balances[c], sib_bal[c]) for c in range(n_currencies)]
        current_hash = poseidon_hash_chip(lay, nxt + [left, right], *pm)
        balances = nxt
    lay.constrain_instance(current_hash, 1)
    for i, b in enumerate(balances):
        lay.constrain_instance(b, 2 + i)
    lay.finish()
    return lay


# ------------------------------------------------------------------ keygen pieces
def compress_selectors(lay: Layouter, max_degree: int = 6) -> List[Dict[int, int]]:
    """halo2 `compress_selectors::process`: returns the new fixed columns (sparse row -> value), in allocation order."""
    out: List[Dict[int, int]] = []
    simple = []
    for s in range(9):
        if SELECTOR_MAX_DEGREE[s] == 0:
            out.append({row: 1 for row in lay.selectors[s]})
        else:
            simple.append(s)
    added = set()
    for i, s in enumerate(simple):
        if s in added:
            continue
        added.add(s)
        d = SELECTOR_MAX_DEGREE[s] - 1
        comb = [s]
        for t in simple[i + 1:]:
            if d + len(comb) == max_degree:
                break
            if t in added:
                continue
            if any(lay.selectors[t] & lay.selectors[u] for u in comb):
                continue
            new_d = max(d, SELECTOR_MAX_DEGREE[t] - 1)
            if new_d + len(comb) + 1 > max_degree:
                continue
            d = new_d
            comb.append(t)
            added.add(t)
        col: Dict[int, int] = {}
        for root, t in enumerate(comb, start=1):
            for row in lay.selectors[t]:
                col[row] = root
        out.append(col)
    return out


def fixed_columns(lay: Layouter) -> List[List[int]]:
    """the 11 fixed columns after selector compression, dense (length n)"""
    cols = [dict(c) for c in lay.fixed] + compress_selectors(lay)
    dense = []
    for c in cols:
        v = [0] * lay.n
        for row, val in c.items():
            v[row] = val
        dense.append(v)
    return dense


def permutation_mapping(lay: Layouter) -> List[List[Tuple[int, int]]]:
    """halo2 permutation keygen `Assembly::copy` replayed over the recorded copy constraints."""
    n, ncols = lay.n, len(PERM_COLUMNS)
    idx = {c: i for i, c in enumerate(PERM_COLUMNS)}
    mapping = [[(c, r) for r in range(n)] for c in range(ncols)]
    aux = [[(c, r) for r in range(n)] for c in range(ncols)]
    sizes = [[1] * n for _ in range(ncols)]
    for lc, lr, rc, rr in lay.copies:
        lcol, rcol = idx[lc], idx[rc]
        left, right = aux[lcol][lr], aux[rcol][rr]
        if left == right:
            continue
        if sizes[left[0]][left[1]] < sizes[right[0]][right[1]]:
            left, right = right, left
        sizes[left[0]][left[1]] += sizes[right[0]][right[1]]
        i = right
        while True:
            aux[i[0]][i[1]] = left
            i = mapping[i[0]][i[1]]
            if i == right:
                break
        mapping[lcol][lr], mapping[rcol][rr] = mapping[rcol][rr], mapping[lcol][lr]
    return mapping


def advice_columns(lay: Layouter) -> List[List[int]]:
    dense = []
    for c in lay.advice:
        v = [0] * lay.n
        for row, val in c.items():
            v[row] = val
        dense.append(v)
    return dense


# ------------------------------------------------------------------ the constraint system itself, from the chip definitions
def constraint_system(n_currencies: int) -> dict:
    """`ConstraintSystem` of `MstInclusionCircuit<_, N_CURRENCIES, _>` after selector compression, in the JSON schema of
    tests/golden/mst_inclusion_cs.json, GENERATED from the chip definitions (not parsed from the 2-currency verifier contract):
    Pow5Chip gates of halo2_gadgets (`full round`, `partial rounds`, `pad-and-add`) for the two Poseidon configurations
    (circuits/merkle_sum_tree.rs:163-186, chips/poseidon/hash.rs:62-72), the `bool constraint`, `swap constraint` and `sum constraint` gates of
    chips/merkle_sum_tree.rs:38-95 -- ONE sum polynomial per currency (:78-88) --, the range check's `lookup_any` (chips/range/range_check.rs:71-84).
    Selector compression assigns: complex lookup selector -> fixed 5; {bool_and_swap, sum, pad_and_add(entry), pad_and_add(middle)}
    -> one column fixed 6 with values 1..4; s_full / s_partial of the two Pow5 configs -> fixed 7..10.
    For N_CURRENCIES = 2 it equals the system parsed from contracts/src/InclusionVerifier.sol gate by gate as polynomials (tests/test_oracle_scale.py)."""
    hx = lambda v: hex(v % R)
    c = lambda v: ["const", hx(v)]
    adv = lambda col, rot=0: ["advice", col, rot]
    fix = lambda col, rot=0: ["fixed", col, rot]
    add = lambda a, b: ["add", a, b]
    mul = lambda a, b: ["mul", a, b]
    neg = lambda a: ["neg", a]
    sub = lambda a, b: add(a, neg(b))

    def pow5(x):
        x2 = mul(x, x)
        return mul(mul(x2, x2), x)
    m, mi = M.MDS, M.MDS_INV

    def pow5_gates(s_full, s_partial, q_pad):
        cur = [add(adv(0), fix(0)), add(adv(1), fix(1))]          # state + round constant (rc_a)
        full = [mul(fix(s_full), sub(add(mul(pow5(cur[0]), c(m[i][0])), mul(pow5(cur[1]), c(m[i][1]))), adv(i, 1))) for i in range(2)]
        mid0 = adv(2)                                               # partial_sbox
        lin = lambda i, rc: add(add(mul(mid0, c(m[i][0])), mul(cur[1], c(m[i][1]))), fix(rc))
        nxt = lambda i: add(mul(adv(0, 1), c(mi[i][0])), mul(adv(1, 1), c(mi[i][1])))
        partial = [mul(fix(s_partial), sub(pow5(cur[0]), mid0)),
                   mul(fix(s_partial), sub(pow5(lin(0, 2)), nxt(0))),
                   mul(fix(s_partial), sub(lin(1, 3), nxt(1)))]
        pad = [mul(q_pad, sub(add(adv(0, -1), adv(0, 0)), adv(0, 1))), mul(q_pad, sub(adv(1, -1), adv(1, 1)))]
        return full + partial + pad

    def combined(v):   # q * prod_{j in 1..4, j != v} (j - q): the gate factor of the selector that got value v in the shared column
        q = fix(6)
        e = q
        for j in range(1, 5):
            if j != v:
                e = mul(e, add(c(j), neg(q)))
        return e
    gates = pow5_gates(7, 8, combined(3)) + pow5_gates(9, 10, combined(4))
    gates.append(mul(mul(combined(1), adv(2)), add(c(1), neg(adv(2)))))                                           # bool
    gates.append(mul(combined(1), sub(add(mul(sub(adv(1), adv(0)), adv(2)), adv(0)), adv(0, 1))))                  # swap (left)
    gates.append(mul(combined(1), sub(add(mul(sub(adv(0), adv(1)), adv(2)), adv(1)), adv(1, 1))))                  # swap (right)
    gates += [mul(combined(2), sub(add(adv(0), adv(1)), adv(2))) for _ in range(n_currencies)]                     # one sum polynomial per currency
    lookup = {"input": [mul(fix(5), add(adv(0), neg(mul(adv(0, 1), c(256)))))], "table": [fix(4)]}
    return {"_source": f"generated by oracle/mst_circuit.py constraint_system({n_currencies}) from the chip definitions",
            "num_advice_columns": 3, "num_fixed_columns": 11, "num_instance_columns": 1, "num_instances": 2 + n_currencies,
            "advice_queries": [[0, 0], [1, 0], [0, 1], [1, 1], [2, 0], [1, -1], [0, -1]],
            "fixed_queries": [[2, 0], [3, 0], [0, 0], [1, 0], [4, 0], [5, 0], [6, 0], [7, 0], [8, 0], [9, 0], [10, 0]],
            "instance_queries": [[0, 0]], "gates": gates, "lookups": [lookup],
            "permutation_columns": [["fixed", 2], ["advice", 0], ["advice", 1], ["fixed", 3], ["advice", 2], ["instance", 0]],
            "degree": 6, "blinding_factors": 5}
