"""Mirror of `zk_prover::merkle_sum_tree::{Entry, Node, MerkleSumTree, MerkleProof, Tree}` over the C ABI.

Reference: zk_prover/src/merkle_sum_tree/{entry.rs, node.rs, mst.rs:74-134, tree.rs:22-186, utils/csv_parser.rs:8-59}.
The tree is built by libsumma_b200 on the GPU (Keccak-256 of usernames, Poseidon leaf / middle hashes, one launch per
level) and stays in HBM; roots, nodes and Merkle proofs are read back on demand.  Field elements are python ints here."""
from __future__ import annotations

import csv
import ctypes
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _lib, fields
from .context import Context, default_context, ptr


@dataclass
class Entry:
    """entry.rs:8-13 (username, balances); the zero entry is `Entry.zero(n)` (entry.rs:30-38)."""
    username: str
    balances: List[int]

    @classmethod
    def zero(cls, n_currencies: int) -> "Entry":
        return cls("0", [0] * n_currencies)


@dataclass
class Node:
    hash: int
    balances: List[int]


@dataclass
class MerkleProof:
    """merkle_sum_tree/mod.rs `MerkleProof`: what `MstInclusionCircuit::init` consumes (circuits/merkle_sum_tree.rs:100-122)."""
    entry_preimage: List[int]                        # [hashed username mod r, balances...]
    root: Node
    sibling_leaf_node_hash_preimage: List[int]
    sibling_middle_node_hash_preimages: List[List[int]]
    path_indices: List[int]


def _fr_list(a: np.ndarray) -> List[int]:
    return [fields.fr_from_mont(a[i]) for i in range(a.shape[0])]


class MerkleSumTree:
    def __init__(self, handle, ctx: Context, entries: Optional[Sequence[Entry]], cryptocurrencies=None, is_sorted=False):
        self._h, self.ctx, self._entries = handle, ctx, entries
        self.cryptocurrencies, self.is_sorted = cryptocurrencies or [], is_sorted
        d, c, ms = ctypes.c_uint32(), ctypes.c_uint32(), ctypes.c_float()
        _lib.check(_lib.lib().sb_mst_shape(self._h, ctypes.byref(d), ctypes.byref(c), ctypes.byref(ms)), "sb_mst_shape")
        self._depth, self.n_currencies, self.build_ms = d.value, c.value, ms.value

    # ---- constructors (mst.rs:74-134) ----
    @classmethod
    def from_entries(cls, entries: Sequence[Entry], cryptocurrencies=None, is_sorted=False, ctx: Optional[Context] = None) -> "MerkleSumTree":
        if len(entries) == 0:
            raise AssertionError("MerkleSumTree: no entries")
        n_cur = len(entries[0].balances)
        names = [e.username.encode() for e in entries]
        for e in entries:
            if len(e.balances) != n_cur:
                raise AssertionError("MerkleSumTree: every entry needs N_CURRENCIES balances")
            if any(not 0 <= b < (1 << 256) for b in e.balances):
                raise AssertionError("balance outside the 256-bit range")
        if all(b < (1 << 64) for e in entries for b in e.balances):
            bal = np.array([[b for b in e.balances] for e in entries], dtype=np.uint64).reshape(len(entries), n_cur)
        else:
            # BigUint balances (entry.rs:10; csv/entry_16_bigints.csv): 32-byte little-endian integers, the wide entry point
            bal = np.frombuffer(b"".join(int(b).to_bytes(32, "little") for e in entries for b in e.balances), dtype=np.uint64).reshape(len(entries), n_cur, 4).copy()
        return cls.from_arrays(names, bal, cryptocurrencies, is_sorted, ctx, entries=list(entries))

    @classmethod
    def from_arrays(cls, names: Sequence[bytes], balances: np.ndarray, cryptocurrencies=None, is_sorted=False, ctx: Optional[Context] = None, entries=None):
        """names: one bytes object per user; balances: (n, N_CURRENCIES) uint64, or (n, N_CURRENCIES, 4) uint64 = 256-bit little-endian
        BigUint balances (reduced mod r like the reference's `big_uint_to_fp`)."""
        ctx = ctx or default_context()
        n = len(names)
        bal = np.ascontiguousarray(balances, dtype=np.uint64)
        wide = bal.ndim == 3
        bal = bal.reshape(n, -1, 4) if wide else bal.reshape(n, -1)
        offs = np.zeros(n + 1, dtype=np.uint32)
        np.cumsum([len(x) for x in names], out=offs[1:])
        blob = np.frombuffer(b"".join(names) + b"\0", dtype=np.uint8).copy()
        h = ctypes.c_void_p()
        fn = _lib.lib().sb_mst_build_wide if wide else _lib.lib().sb_mst_build
        _lib.check(fn(ctx.handle, ptr(blob), ptr(offs), ptr(bal), ctypes.c_size_t(n), ctypes.c_uint32(bal.shape[1]), ctypes.byref(h)), "sb_mst_build")
        return cls(h, ctx, entries, cryptocurrencies, is_sorted)

    @classmethod
    def from_csv(cls, path: str, ctx: Optional[Context] = None, sort: bool = False) -> "MerkleSumTree":
        """csv_parser.rs:8-59: header `username,balance_<name>_<chain>,...`; `sort=True` is `from_csv_sorted` (mst.rs:87-94)."""
        with open(path) as f:
            rows = list(csv.reader(f))
        cur = []
        for h in rows[0][1:]:
            parts = h.split("_")
            if len(parts) != 3 or parts[0] != "balance":
                raise ValueError(f"Invalid header: {h}")
            cur.append({"name": parts[1], "chain": parts[2]})
        entries = [Entry(r[0], [int(x) for x in r[1:]]) for r in rows[1:] if r]  # empty lines are skipped like the csv crate does
        if sort:
            entries.sort(key=lambda e: e.username)
        return cls.from_entries(entries, cur, sort, ctx)

    @classmethod
    def from_leaf_preimages(cls, preimages: np.ndarray, n_currencies: int, ctx: Optional[Context] = None) -> "MerkleSumTree":
        """build_merkle_tree_from_leaves (utils/build_tree.rs:5-28): (2^depth, N_CURRENCIES + 1, 4) uint64 Montgomery preimages."""
        ctx = ctx or default_context()
        pre = np.ascontiguousarray(preimages, dtype=np.uint64).reshape(-1, n_currencies + 1, 4)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_mst_build_from_preimages(ctx.handle, ptr(pre), ctypes.c_size_t(pre.shape[0]), ctypes.c_uint32(n_currencies), ctypes.byref(h)),
                   "sb_mst_build_from_preimages")
        return cls(h, ctx, None)

    # ---- Tree trait (tree.rs:7-21) ----
    def depth(self) -> int:
        return self._depth

    def node(self, level: int, index: int) -> Node:
        hs = np.zeros(4, dtype=np.uint64)
        bl = np.zeros((self.n_currencies, 4), dtype=np.uint64)
        _lib.check(_lib.lib().sb_mst_node(self._h, ctypes.c_uint32(level), ctypes.c_size_t(index), ptr(hs), ptr(bl)), "sb_mst_node")
        return Node(fields.fr_from_mont(hs), _fr_list(bl))

    def root(self) -> Node:
        return self.node(self._depth, 0)

    def level_hashes(self, level: int) -> np.ndarray:
        out = np.zeros((1 << (self._depth - level), 4), dtype=np.uint64)
        _lib.check(_lib.lib().sb_mst_level_hashes(self._h, ctypes.c_uint32(level), ptr(out)), "sb_mst_level_hashes")
        return out

    def get_entry(self, index: int) -> Entry:
        if self._entries is None:
            raise AssertionError("tree built from preimages keeps no entries")
        return self._entries[index] if index < len(self._entries) else Entry.zero(self.n_currencies)

    def generate_proofs(self, indices: Sequence[int]) -> List[MerkleProof]:
        """tree.rs:85-137 for many users in one launch."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        m, d, c = idx.shape[0], self._depth, self.n_currencies
        per = 2 * (c + 1) + max(d - 1, 0) * (c + 2)
        pre = np.zeros((m, per, 4), dtype=np.uint64)
        path = np.zeros((m, max(d, 1)), dtype=np.uint8)
        _lib.check(_lib.lib().sb_mst_proofs(self._h, ptr(idx), ctypes.c_size_t(m), ptr(pre), ptr(path)), "sb_mst_proofs")
        root = self.root()
        out = []
        for j in range(m):
            vals = _fr_list(pre[j])
            mids = [vals[2 * (c + 1) + t * (c + 2): 2 * (c + 1) + (t + 1) * (c + 2)] for t in range(max(d - 1, 0))]
            out.append(MerkleProof(vals[: c + 1], root, vals[c + 1: 2 * (c + 1)], mids, [int(x) for x in path[j, :d]]))
        return out

    def raw_proofs(self, indices: Sequence[int]):
        """`generate_proofs` without the python objects: (preimages (m, per, 4) uint64 Montgomery, path_indices (m, depth) uint8), the layout
        sb_mst_proofs writes and sb_mst_inclusion_witness reads"""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        m, d, c = idx.shape[0], self._depth, self.n_currencies
        per = 2 * (c + 1) + max(d - 1, 0) * (c + 2)
        pre = np.zeros((m, per, 4), dtype=np.uint64)
        path = np.zeros((m, max(d, 1)), dtype=np.uint8)
        _lib.check(_lib.lib().sb_mst_proofs(self._h, ptr(idx), ctypes.c_size_t(m), ptr(pre), ptr(path)), "sb_mst_proofs")
        return pre, path

    def index_of_username(self, username: str) -> int:
        """mst.rs:200-216: linear search, or binary search when the tree was built sorted"""
        if self._entries is None:
            raise AssertionError("tree built from preimages keeps no entries")
        if self.is_sorted:
            import bisect
            names = [e.username for e in self._entries]
            i = bisect.bisect_left(names, username)
            if i < len(names) and names[i] == username:
                return i
        else:
            for i, e in enumerate(self._entries):
                if e.username == username:
                    return i
        raise KeyError("Username not found")

    def update_leaf(self, username: str, new_balances: Sequence[int]) -> Node:
        """mst.rs:158-197: store the new balances, rehash the leaf and its path on the GPU, return the new root"""
        index = self.index_of_username(username)
        if len(new_balances) != self.n_currencies:
            raise AssertionError("update_leaf: N_CURRENCIES balances expected")
        hs = np.zeros(4, dtype=np.uint64)
        bl = np.zeros((self.n_currencies, 4), dtype=np.uint64)
        if all(0 <= int(b) < (1 << 64) for b in new_balances):
            bal = np.array([int(b) for b in new_balances], dtype=np.uint64)
            _lib.check(_lib.lib().sb_mst_update_leaf(self._h, ctypes.c_size_t(index), ptr(bal), ptr(hs), ptr(bl)), "sb_mst_update_leaf")
        else:
            bal = np.frombuffer(b"".join(int(b).to_bytes(32, "little") for b in new_balances), dtype=np.uint64).copy()
            _lib.check(_lib.lib().sb_mst_update_leaf_wide(self._h, ctypes.c_size_t(index), ptr(bal), ptr(hs), ptr(bl)), "sb_mst_update_leaf_wide")
        self._entries[index] = Entry(username, [int(x) for x in new_balances])
        return Node(fields.fr_from_mont(hs), _fr_list(bl))

    def verify_proofs(self, proofs: Sequence[MerkleProof]) -> List[bool]:
        """tree.rs:139-186 on the GPU, one thread per proof"""
        if not proofs:
            return []
        c, d = self.n_currencies, len(proofs[0].path_indices)
        per = 2 * (c + 1) + max(d - 1, 0) * (c + 2)
        pre = np.zeros((len(proofs), per, 4), dtype=np.uint64)
        path = np.zeros((len(proofs), max(d, 1)), dtype=np.uint8)
        for j, p in enumerate(proofs):
            vals = list(p.entry_preimage) + list(p.sibling_leaf_node_hash_preimage) + [x for m in p.sibling_middle_node_hash_preimages for x in m]
            if len(vals) != per or len(p.path_indices) != d:
                raise AssertionError("verify_proofs: proofs of different shapes")
            for i, v in enumerate(vals):
                pre[j, i] = fields.fr_to_mont(v)
            path[j, :d] = p.path_indices
        root = proofs[0].root
        rh = fields.fr_to_mont(root.hash)
        rb = np.stack([fields.fr_to_mont(b) for b in root.balances])
        ok = np.zeros(len(proofs), dtype=np.uint8)
        _lib.check(_lib.lib().sb_mst_verify_proofs(self.ctx.handle, ctypes.c_uint32(c), ctypes.c_uint32(d), ptr(pre), ptr(path), ptr(rh), ptr(rb), ctypes.c_size_t(len(proofs)), ptr(ok)),
                   "sb_mst_verify_proofs")
        return [bool(x) for x in ok]

    def verify_proof(self, proof: MerkleProof) -> bool:
        return self.verify_proofs([proof])[0]

    def generate_proof(self, index: int) -> MerkleProof:
        if not 0 <= index < (1 << self._depth):
            raise IndexError("Index out of bounds")
        return self.generate_proofs([index])[0]

    def close(self):
        if self._h:
            _lib.lib().sb_mst_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
