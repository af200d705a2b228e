"""CPU ORACLE (test infrastructure) -- Summa's Merkle sum tree and its Poseidon hash, restated.

Follows zk_prover/src/merkle_sum_tree/{entry.rs:15-38, node.rs:16-85, mst.rs:74-134, tree.rs:85-137,
utils/build_tree.rs:5-78, utils/csv_parser.rs:8-59, utils/operation_helpers.rs:10-12} and the
halo2_gadgets Poseidon primitive (un-vendored) with the constants of
chips/poseidon/poseidon_params.rs (tests/golden/poseidon_params.json).
Pinned by the Rust tests' known answers: leaf hashes circuits/tests.rs:341,346, root hash
backend/src/tests.rs:265, root balances merkle_sum_tree/tests.rs:24  (tests/test_oracle_circuit.py)."""
from __future__ import annotations

import csv
import json
import os
from typing import List, Tuple

from . import bn254 as B
from .keccak import keccak256

R = B.R
_GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
_P = json.load(open(os.path.join(_GOLDEN, "poseidon_params.json")))
RC = [[int(x, 16) for x in row] for row in _P["round_constants"]]
MDS = [[int(x, 16) for x in row] for row in _P["mds"]]
MDS_INV = [[int(x, 16) for x in row] for row in _P["mds_inv"]]
R_F, R_P = 8, 56


def mds_mul(s):
    return [(MDS[0][0] * s[0] + MDS[0][1] * s[1]) % R, (MDS[1][0] * s[0] + MDS[1][1] * s[1]) % R]


def permute(s: List[int]) -> List[int]:
    s = list(s)
    for r in range(R_F + R_P):
        if r < R_F // 2 or r >= R_F // 2 + R_P:
            s = mds_mul([pow((s[0] + RC[r][0]) % R, 5, R), pow((s[1] + RC[r][1]) % R, 5, R)])
        else:
            s = mds_mul([pow((s[0] + RC[r][0]) % R, 5, R), (s[1] + RC[r][1]) % R])
    return s


def poseidon_hash(inputs: List[int]) -> int:
    """ConstantLength<L>, WIDTH 2, RATE 1: state [0, L * 2^64]; absorb one element per permutation."""
    s = [0, (len(inputs) << 64) % R]
    for x in inputs:
        s[0] = (s[0] + x) % R
        s = permute(s)
    return s[0]


class Entry:
    def __init__(self, username: str, balances: List[int], zero: bool = False):
        self.username = username
        self.balances = list(balances)
        self.hashed_username = 0 if zero else int.from_bytes(keccak256(username.encode()), "big")

    def preimage(self) -> List[int]:
        return [self.hashed_username % R] + [b % R for b in self.balances]


class MerkleSumTree:
    def __init__(self, entries: List[Entry]):
        n_cur = len(entries[0].balances)
        depth = max(0, (len(entries) - 1).bit_length())
        entries = list(entries) + [Entry("0", [0] * n_cur, zero=True) for _ in range((1 << depth) - len(entries))]
        self.entries, self.depth, self.n_currencies = entries, depth, n_cur
        level = [(poseidon_hash(e.preimage()), [b % R for b in e.balances]) for e in entries]
        self.nodes = [level]
        for _ in range(depth):
            nxt = []
            for i in range(0, len(level), 2):
                bal = [(a + b) % R for a, b in zip(level[i][1], level[i + 1][1])]
                nxt.append((poseidon_hash(bal + [level[i][0], level[i + 1][0]]), bal))
            level = nxt
            self.nodes.append(level)
        self.root = level[0]

    @classmethod
    def from_csv(cls, path: str) -> "MerkleSumTree":
        with open(path) as f:
            rows = list(csv.reader(f))
        return cls([Entry(r[0], [int(x) for x in r[1:]]) for r in rows[1:] if r])  # the csv crate skips empty lines (csv/entry_17.csv ends with one)

    def generate_proof(self, index: int) -> dict:
        sib = index + 1 if index % 2 == 0 else index - 1
        path, mids, cur = [], [], index
        for level in range(self.depth):
            pos = cur % 2
            sidx = cur - pos + (1 - pos)
            if level > 0:
                l, r = self.nodes[level - 1][2 * sidx], self.nodes[level - 1][2 * sidx + 1]
                mids.append([(a + b) % R for a, b in zip(l[1], r[1])] + [l[0], r[0]])
            path.append(pos)
            cur //= 2
        return {"entry": self.entries[index], "root": self.root, "sibling_leaf_node_hash_preimage": self.entries[sib].preimage(),
                "sibling_middle_node_hash_preimages": mids, "path_indices": path}
