"""Writes tests/golden/mst_inclusion_assignment.npz: the keygen + witness output of the reference circuit
`MstInclusionCircuit<4, 2, 8>` for csv/entry_16.csv, user 0, in SPARSE form (the circuit uses < 1500 rows
whatever k is), produced by the oracle's restatement of the circuit (oracle/mst_circuit.py, pinned
on the reference vk).  bench.py and the k >= 17 GPU tests expand it to any k without importing oracle/.

    python tests/golden/make_assignment.py
"""
import os
import sys

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import bn254 as B  # noqa: E402
from oracle import mst as M  # noqa: E402
from oracle import mst_circuit as C  # noqa: E402

tree = M.MerkleSumTree.from_csv(os.path.join(here, "entry_16.csv"))
k = 11
lay = C.synthesize(k, tree.generate_proof(0), 4, 2, 8)
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
instances = [tree.nodes[0][0][0], tree.root[0]] + tree.root[1]
np.savez_compressed(os.path.join(here, "mst_inclusion_assignment.npz"),
                    fixed_cells=np.array(fixed_cells, dtype=np.uint32), fixed_values=np.stack(fixed_vals),
                    perm_cells=np.array(perm_cells, dtype=np.uint32),
                    advice_cells=np.array(adv_cells, dtype=np.uint32), advice_values=np.stack(adv_vals),
                    instances=np.stack([mont(v) for v in instances]), rows_used=np.array([max(lay.next_free.values())]))
print("fixed cells", len(fixed_cells), "perm cells", len(perm_cells), "advice cells", len(adv_cells))
