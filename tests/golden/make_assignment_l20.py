"""Writes tests/golden/mst_inclusion_assignment_l20.npz: keygen + witness of `MstInclusionCircuit<20, 2, 8>` (BASELINE configs[4]: a tree of
2^20 users, LEVELS = 20, minimum k = 13) in the same sparse form as mst_inclusion_assignment.npz.  A 2^20-leaf tree is not needed for
one Merkle path: the path's sibling nodes are fabricated (seeded), the root is whatever they hash up to, and the circuit's constraints
(Poseidon hashes, sums, swaps, 8-byte range checks at every level) are all satisfied.  Produced by the oracle's restatement of the circuit.

    python tests/golden/make_assignment_l20.py
"""
import os
import random
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import bn254 as B  # noqa: E402
from oracle import mst as M  # noqa: E402
from oracle import mst_circuit as C  # noqa: E402

LEVELS, N_CUR, K = 20, 2, 13
rnd = random.Random(20)
entry = M.Entry("user_123456", [1000, 2000])
sib_leaf = [rnd.randrange(B.R), 5, 7]
index = 123456
path = [(index >> l) & 1 for l in range(LEVELS)]
mids = []
node_hash, bal = M.poseidon_hash(entry.preimage()), list(entry.balances)
sib_hash, sib_bal = M.poseidon_hash(sib_leaf), sib_leaf[1:]
for level in range(LEVELS):
    if level > 0:
        pre = [level * 11, level * 13, rnd.randrange(B.R), rnd.randrange(B.R)]   # [balances..., hash_l, hash_r] of the sibling
        mids.append(pre)
        sib_hash, sib_bal = M.poseidon_hash(pre), pre[:N_CUR]
    bal = [a + b for a, b in zip(bal, sib_bal)]
    hs = [node_hash, sib_hash] if path[level] == 0 else [sib_hash, node_hash]
    node_hash = M.poseidon_hash(bal + hs)
proof = {"entry": entry, "root": (node_hash, bal), "sibling_leaf_node_hash_preimage": sib_leaf, "sibling_middle_node_hash_preimages": mids, "path_indices": path}
lay = C.synthesize(K, proof, LEVELS, N_CUR, 8)
mont = lambda x: np.frombuffer(B.fr_to_mont_bytes(x), dtype=np.uint64)
fixed_cells, fixed_vals = [], []
for col, dense in enumerate(C.fixed_columns(lay)):
    for row, v in enumerate(dense):
        if v:
            fixed_cells.append((col, row))
            fixed_vals.append(mont(v))
perm_cells = []
for col, rows in enumerate(C.permutation_mapping(lay)):
    for row, (tc, tr) in enumerate(rows):
        if (tc, tr) != (col, row):
            perm_cells.append((col, row, tc, tr))
adv_cells, adv_vals = [], []
for col, dense in enumerate(C.advice_columns(lay)):
    for row, v in enumerate(dense):
        if v:
            adv_cells.append((col, row))
            adv_vals.append(mont(v))
instances = [M.poseidon_hash(entry.preimage()), node_hash] + bal
np.savez_compressed(os.path.join(here, "mst_inclusion_assignment_l20.npz"),
                    fixed_cells=np.array(fixed_cells, dtype=np.uint32), fixed_values=np.stack(fixed_vals),
                    perm_cells=np.array(perm_cells, dtype=np.uint32),
                    advice_cells=np.array(adv_cells, dtype=np.uint32), advice_values=np.stack(adv_vals),
                    instances=np.stack([mont(v) for v in instances]), rows_used=np.array([max(lay.next_free.values())]))
print("fixed cells", len(fixed_cells), "perm cells", len(perm_cells), "advice cells", len(adv_cells), "rows used", max(lay.next_free.values()))
