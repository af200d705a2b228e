"""Writes the CPU-oracle golden proofs the GPU prover must reproduce byte for byte at BASELINE's K values:

  golden_proof_k17.npz   MstInclusionCircuit<4,2,8> (csv/entry_16.csv, user 0; witness = mst_inclusion_assignment.npz) at k = 17
  golden_proof_k20.npz   MstInclusionCircuit<20,2,8> at k = 20 for user 123456 of a 2^20-user tree (users `user_{i}`, balances uniform
                         in [0, 2^40), numpy default_rng(20): SURVEY 8d config 3), tree built by the C oracle
  mst_inclusion_assignment_l20_tree.npz   that circuit's keygen + witness in sparse form, the Merkle path, the tree's root
  mst_inclusion_assignment_l23_n8_tree.npz + mst_inclusion_cs_n8.json + golden_proof_k15_n8.npz   the same for BASELINE configs[3]'s
                         MstInclusionCircuit<23,8,8> (2^23 users, 8 currencies; constraint system generated from the chip definitions)

Each proof file holds: proof (Keccak / EVM transcript, 2144 B), instances, the 11 fixed + 6 permutation commitments of the key
(keygen_vk's output for this k and SRS), tau of the unsafe SRS, the ChaCha20 seed and vk.transcript_repr used.
Everything is computed by oracle/ (python + oracle/halo2_cpu.c); ~25 s for k = 17 and ~5 min for k = 20 on 8 cores.

    python tests/golden/make_golden_proofs.py [17] [20] [23]
"""
import json
import os
import sys
import time

import numpy as np

here = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(here)))
from oracle import bn254 as B  # noqa: E402
from oracle import cpu  # noqa: E402
from oracle import halo2_prover as HP  # noqa: E402
from oracle import mst as M  # noqa: E402
from oracle import mst_circuit as C  # noqa: E402
from oracle.chacha import ChaCha20Rng  # noqa: E402
from oracle.transcript import KeccakTranscript  # noqa: E402

THREADS = os.cpu_count() or 1
SEED_U64, TRANSCRIPT_REPR = 42, 0x1234
um = lambda x: B.fr_from_mont_bytes(np.ascontiguousarray(x).tobytes())
mont = lambda x: np.frombuffer(B.fr_to_mont_bytes(x), dtype=np.uint64)


def tau_for(k):
    return 0x5A110000 + k


def prove(k, fx, tag, cs_file="mst_inclusion_cs.json", out=None):
    cs = json.load(open(os.path.join(here, cs_file)))
    t0 = time.time()
    params = HP.Params.setup(k, tau_for(k), THREADS)
    pk = HP.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], TRANSCRIPT_REPR)
    n = 1 << k
    adv = np.zeros((3, n, 4), dtype=np.uint64)
    c = fx["advice_cells"]
    adv[c[:, 0], c[:, 1]] = fx["advice_values"]
    inst = [um(v) for v in fx["instances"]]
    t1 = time.time()
    tr = KeccakTranscript()
    HP.create_proof(params, pk, inst, adv, ChaCha20Rng.seed_from_u64(SEED_U64), tr)
    proof = tr.finalize()
    t2 = time.time()
    g1 = lambda p: np.frombuffer(B.g1_to_mont_bytes(p), dtype=np.uint64)
    np.savez_compressed(os.path.join(here, out or f"golden_proof_k{k}.npz"), proof=np.frombuffer(proof, dtype=np.uint8), instances=fx["instances"],
                        fixed_comms=np.stack([g1(p) for p in pk.fixed_commitments]), sigma_comms=np.stack([g1(p) for p in pk.sigma_commitments]),
                        k=np.array([k]), tau=np.array([tau_for(k)], dtype=np.uint64), seed_u64=np.array([SEED_U64]), transcript_repr=np.array([TRANSCRIPT_REPR]),
                        witness=np.array([tag]))
    print(f"k={k}: setup+keygen {t1 - t0:.1f} s, create_proof {t2 - t1:.1f} s ({THREADS} threads), proof {len(proof)} B")


def make_tree_assignment(LEVELS, N_CUR, INDEX, seed, out_name, k_min):
    n_users = 1 << LEVELS
    bal = np.random.default_rng(seed).integers(0, 1 << 40, size=(n_users, N_CUR), dtype=np.uint64)
    names = [b"user_%d" % i for i in range(n_users)]
    t0 = time.time()
    tree = cpu.MstC(names, bal)
    print(f"2^{LEVELS}-user, {N_CUR}-currency tree on the CPU oracle: {time.time() - t0:.1f} s ({THREADS} threads)")
    # Tree::generate_proof (tree.rs:85-137) from the flat arrays
    sib = INDEX ^ 1
    entry = M.Entry(names[INDEX].decode(), [int(x) for x in bal[INDEX]])
    assert entry.hashed_username % B.R == um(tree.unames[INDEX])
    sib_pre = [um(tree.unames[sib])] + [int(x) for x in bal[sib]]
    path, mids, cur = [], [], INDEX
    for level in range(LEVELS):
        pos = cur & 1
        sidx = cur ^ 1
        if level > 0:
            (hl, bl), (hr, br) = tree.node(level - 1, 2 * sidx), tree.node(level - 1, 2 * sidx + 1)
            mids.append([(um(a) + um(b)) % B.R for a, b in zip(bl, br)] + [um(hl), um(hr)])
        path.append(pos)
        cur >>= 1
    rh, rb = tree.root()
    root = (um(rh), [um(x) for x in rb])
    assert root[1] == [int(bal[:, c].astype(object).sum()) for c in range(N_CUR)]
    proof = {"entry": entry, "root": root, "sibling_leaf_node_hash_preimage": sib_pre, "sibling_middle_node_hash_preimages": mids, "path_indices": path}
    lay = C.synthesize(k_min, proof, LEVELS, N_CUR, 8)   # rows used do not depend on k; k_min is the circuit's minimum
    fixed_cells, fixed_vals, perm_cells, adv_cells, adv_vals = [], [], [], [], []
    for col, dense in enumerate(C.fixed_columns(lay)):
        for row, v in enumerate(dense):
            if v:
                fixed_cells.append((col, row)); fixed_vals.append(mont(v))
    for col, rows in enumerate(C.permutation_mapping(lay)):
        for row, (tc, tr) in enumerate(rows):
            if (tc, tr) != (col, row):
                perm_cells.append((col, row, tc, tr))
    for col, dense in enumerate(C.advice_columns(lay)):
        for row, v in enumerate(dense):
            if v:
                adv_cells.append((col, row)); adv_vals.append(mont(v))
    leaf_hash = M.poseidon_hash(entry.preimage())
    assert leaf_hash == um(tree.node(0, INDEX)[0])
    instances = [leaf_hash, root[0]] + root[1]
    out = os.path.join(here, out_name)
    np.savez_compressed(out, fixed_cells=np.array(fixed_cells, dtype=np.uint32), fixed_values=np.stack(fixed_vals), perm_cells=np.array(perm_cells, dtype=np.uint32),
                        advice_cells=np.array(adv_cells, dtype=np.uint32), advice_values=np.stack(adv_vals), instances=np.stack([mont(v) for v in instances]),
                        rows_used=np.array([max(lay.next_free.values())]), user_index=np.array([INDEX]), n_users=np.array([n_users]), balance_seed=np.array([seed]),
                        n_currencies=np.array([N_CUR]), levels=np.array([LEVELS]),
                        root_hash=mont(root[0]), root_balances=np.stack([mont(v) for v in root[1]]),
                        path_indices=np.array(path, dtype=np.uint8), sibling_leaf_preimage=np.stack([mont(v) for v in sib_pre]),
                        sibling_middle_preimages=np.stack([np.stack([mont(v) for v in pre]) for pre in mids]))
    print("wrote", out, "rows used", max(lay.next_free.values()))


if __name__ == "__main__":
    cpu.set_threads(THREADS)
    which = [int(x) for x in sys.argv[1:]] or [17, 20]
    if 17 in which:
        prove(17, np.load(os.path.join(here, "mst_inclusion_assignment.npz")), "MstInclusionCircuit<4,2,8>, entry_16.csv user 0")
    if 20 in which:
        l20 = os.path.join(here, "mst_inclusion_assignment_l20_tree.npz")
        if not os.path.exists(l20):
            make_tree_assignment(20, 2, 123456, 20, "mst_inclusion_assignment_l20_tree.npz", 13)
        prove(20, np.load(l20), "MstInclusionCircuit<20,2,8>, user 123456 of the 2^20-user tree (default_rng(20) balances)")
    if 23 in which:
        # BASELINE configs[3]: MstInclusionCircuit<23,8,8> over a 2^23-user, 8-currency tree; the constraint system comes from the chip definitions
        # (oracle/mst_circuit.py constraint_system(8) -> mst_inclusion_cs_n8.json).  A k = 23 CPU proof would take ~20 minutes and 60 GB, so the
        # golden proof of THIS circuit is made at its minimum k = 15 (same witness, same key cells); the k = 23 GPU proof is judged by
        # oracle/halo2_verifier.py and the closed-form key check.
        json.dump(C.constraint_system(8), open(os.path.join(here, "mst_inclusion_cs_n8.json"), "w"))
        l23 = os.path.join(here, "mst_inclusion_assignment_l23_n8_tree.npz")
        if not os.path.exists(l23):
            make_tree_assignment(23, 8, 7654321, 23, "mst_inclusion_assignment_l23_n8_tree.npz", 15)
        prove(15, np.load(l23), "MstInclusionCircuit<23,8,8>, user 7654321 of the 2^23-user 8-currency tree (default_rng(23) balances), at k = 15",
              cs_file="mst_inclusion_cs_n8.json", out="golden_proof_k15_n8.npz")
