"""GPU parity tests of the Merkle-sum-tree build (python -m pytest tests -m gpu).

The CUDA tree (Keccak-256 usernames, Poseidon leaves / middle nodes, Merkle proofs) must equal the CPU oracle
(oracle/mst.py, pinned on the Rust tests' known answers) bit for bit; at 2^20 users the size-independent properties
are root balances = column sums and oracle-side verification of GPU Merkle proofs (tree.rs:139-186)."""
import json
import os

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import mst as M

pytestmark = pytest.mark.gpu


def oracle_verify(proof, n_cur):
    """Tree::verify_proof (tree.rs:139-186) with the oracle's Poseidon."""
    node_hash = M.poseidon_hash(proof.entry_preimage)
    bal = list(proof.entry_preimage[1:])
    sib_hash, sib_bal = M.poseidon_hash(proof.sibling_leaf_node_hash_preimage), proof.sibling_leaf_node_hash_preimage[1:]
    for lvl, pos in enumerate(proof.path_indices):
        if lvl > 0:
            pre = proof.sibling_middle_node_hash_preimages[lvl - 1]
            sib_hash, sib_bal = M.poseidon_hash(pre), pre[:n_cur]
        bal = [(a + b) % B.R for a, b in zip(bal, sib_bal)]
        hs = [node_hash, sib_hash] if pos == 0 else [sib_hash, node_hash]
        node_hash = M.poseidon_hash(bal + hs)
    return node_hash == proof.root.hash and bal == proof.root.balances


def test_entry_16_csv_known_answers(ctx, golden_dir):
    import circuits_halo2_b200 as sb
    gold = json.load(open(os.path.join(golden_dir, "mst_hashes.json")))
    t = sb.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"), ctx)
    assert t.depth() == 4 and t.n_currencies == 2
    assert hex(t.node(0, 0).hash) in gold["circuit_tests_hex"]      # circuits/tests.rs:341
    assert hex(t.node(0, 1).hash) in gold["circuit_tests_hex"]      # circuits/tests.rs:346
    root = t.root()
    assert hex(root.hash) in gold["backend_tests_hex"]              # backend/src/tests.rs:265
    assert root.balances == [556862, 556862]                        # merkle_sum_tree/tests.rs:24
    o = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    for level in range(5):
        for i in range(1 << (4 - level)):
            n = t.node(level, i)
            assert (n.hash, n.balances) == (o.nodes[level][i][0], o.nodes[level][i][1]), (level, i)
    for idx in (0, 1, 7, 15):
        p, q = t.generate_proof(idx), o.generate_proof(idx)
        assert p.entry_preimage == q["entry"].preimage()
        assert p.sibling_leaf_node_hash_preimage == q["sibling_leaf_node_hash_preimage"]
        assert p.sibling_middle_node_hash_preimages == q["sibling_middle_node_hash_preimages"]
        assert p.path_indices == q["path_indices"]
        assert oracle_verify(p, 2)
    with pytest.raises(IndexError):
        t.generate_proof(16)


@pytest.mark.parametrize("n_entries,n_cur", [(1, 1), (2, 2), (3, 2), (5, 3), (16, 1), (100, 2), (513, 4), (1000, 8)])
def test_ragged_sizes_match_oracle(ctx, n_entries, n_cur):
    """non-power-of-two entry counts are padded with zero entries (mst.rs:106-114); long and empty usernames exercise the Keccak padding"""
    import circuits_halo2_b200 as sb
    rng = np.random.default_rng(n_entries * 10 + n_cur)
    names = []
    for i in range(n_entries):
        ln = [0, 1, 8, 135, 136, 137, 271, 272, 300][i % 9] if i % 5 == 0 else int(rng.integers(1, 24))
        names.append(bytes(rng.integers(33, 127, size=ln, dtype=np.uint8)).decode())
    bal = rng.integers(0, 1 << 63, size=(n_entries, n_cur), dtype=np.uint64)
    bal[0, :] = np.uint64((1 << 64) - 1)
    ents = [sb.Entry(nm, [int(x) for x in bal[i]]) for i, nm in enumerate(names)]
    t = sb.MerkleSumTree.from_entries(ents, ctx=ctx)
    o = M.MerkleSumTree([M.Entry(nm, [int(x) for x in bal[i]]) for i, nm in enumerate(names)])
    assert t.depth() == o.depth
    for level in range(o.depth + 1):
        got = t.level_hashes(level)
        ref = np.stack([np.frombuffer(B.fr_to_mont_bytes(h), dtype=np.uint64) for h, _ in o.nodes[level]])
        assert (got == ref).all(), level
    r = t.root()
    assert (r.hash, r.balances) == (o.root[0], o.root[1])
    if o.depth > 0:
        idx = [0, n_entries - 1, (1 << o.depth) - 1]
        for p, i in zip(t.generate_proofs(idx), idx):
            q = o.generate_proof(i)
            assert p.sibling_middle_node_hash_preimages == q["sibling_middle_node_hash_preimages"] and p.path_indices == q["path_indices"]
            assert p.sibling_leaf_node_hash_preimage == q["sibling_leaf_node_hash_preimage"]


def test_from_leaf_preimages_equals_from_entries(ctx):
    import circuits_halo2_b200 as sb
    ents = [sb.Entry(f"user_{i}", [i * 7 + 1, i * 11 + 2]) for i in range(64)]
    t = sb.MerkleSumTree.from_entries(ents, ctx=ctx)
    pre = np.stack([np.stack([np.frombuffer(B.fr_to_mont_bytes(v), dtype=np.uint64) for v in M.Entry(e.username, e.balances).preimage()]) for e in ents])
    u = sb.MerkleSumTree.from_leaf_preimages(pre, 2, ctx)
    assert (t.level_hashes(0) == u.level_hashes(0)).all() and t.root() == u.root()
    with pytest.raises(Exception):
        sb.MerkleSumTree.from_leaf_preimages(pre[:48], 2, ctx)   # build_tree.rs:17: the leaf layer must be a power of two


def test_two_to_the_twenty_users(ctx):
    """config 3's snapshot: 2^20 users, 2 currencies.  Properties: root balances are the column sums; GPU Merkle proofs verify with the oracle's Poseidon."""
    import circuits_halo2_b200 as sb
    n = 1 << 20
    rng = np.random.default_rng(20)
    bal = rng.integers(0, 1 << 40, size=(n, 2), dtype=np.uint64)
    names = [b"user_%d" % i for i in range(n)]
    t = sb.MerkleSumTree.from_arrays(names, bal, ctx=ctx)
    assert t.depth() == 20
    root = t.root()
    assert root.balances == [int(bal[:, 0].astype(object).sum()), int(bal[:, 1].astype(object).sum())]
    idx = [0, 1, 123456, n - 1]
    for p, i in zip(t.generate_proofs(idx), idx):
        assert p.entry_preimage == M.Entry(names[i].decode(), [int(x) for x in bal[i]]).preimage()
        assert oracle_verify(p, 2)
    print(f"\n[mst] 2^20 users x 2 currencies built in {t.build_ms:.1f} ms on the device")


def test_verify_update_and_sorted_like_the_reference_tests(ctx, golden_dir):
    """merkle_sum_tree/tests.rs: test_mst (verify, tampered proofs rejected), test_update_mst_leaf, test_update_invalid_mst_leaf, test_sorted_mst."""
    import copy
    import circuits_halo2_b200 as sb
    path = os.path.join(golden_dir, "entry_16.csv")
    t = sb.MerkleSumTree.from_csv(path, ctx)
    proofs = t.generate_proofs(list(range(16)))
    assert all(t.verify_proofs(proofs)) and all(oracle_verify(p, 2) for p in proofs)
    bad1 = copy.deepcopy(proofs[0])
    bad1.sibling_leaf_node_hash_preimage[1] += 1            # tests.rs:52-60: a wrong sibling balance
    bad2 = copy.deepcopy(proofs[0])
    bad2.path_indices[0] ^= 1                               # tests.rs:62-65: a wrong path index
    bad3 = copy.deepcopy(proofs[3])
    bad3.entry_preimage[0] = (bad3.entry_preimage[0] + 1) % B.R
    assert t.verify_proofs([bad1, bad2, bad3, proofs[5]]) == [False, False, False, True]
    # update_leaf: change, compare with the oracle tree of the changed csv, change back (tests.rs:69-92)
    root0 = t.root()
    name = t.get_entry(7).username
    old = list(t.get_entry(7).balances)
    r1 = t.update_leaf(name, [11, 22])
    o = M.MerkleSumTree.from_csv(path)
    ents = [M.Entry(e.username, list(e.balances)) for e in o.entries]
    ents[7] = M.Entry(name, [11, 22])
    o2 = M.MerkleSumTree(ents)
    assert (r1.hash, r1.balances) == (o2.root[0], o2.root[1]) and r1.hash != root0.hash
    for level in range(5):
        got = t.level_hashes(level)
        assert [fr_int(x) for x in got] == [h for h, _ in o2.nodes[level]], level
    assert t.verify_proof(t.generate_proof(7))
    r2 = t.update_leaf(name, old)
    assert (r2.hash, r2.balances) == (root0.hash, root0.balances)
    with pytest.raises(KeyError):
        t.update_leaf("non_existing_user", [1, 2])           # tests.rs:95-106
    # sorted tree (tests.rs:110-127): same balances, different root hash, entries ordered by username
    ts = sb.MerkleSumTree.from_csv(path, ctx, sort=True)
    assert ts.root().balances == root0.balances and ts.root().hash != root0.hash
    names = [ts.get_entry(i).username for i in range(16)]
    assert names == sorted(names) and ts.index_of_username(names[5]) == 5


def fr_int(limbs):
    from circuits_halo2_b200 import fields
    return fields.fr_from_mont(limbs)


def test_entry_13_and_17_csv_goldens_with_zero_padding(ctx, golden_dir):
    """merkle_sum_tree/tests.rs:205-264 (`test_tree_with_zero_element_1/2`): 13 entries pad to 16 (depth 4), 17 pad to 32 (depth 5) with zero entries;
    root balances are the reference's goldens; every index proves and verifies, the first out-of-range index errors."""
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200._lib import SummaB200Error
    for name, n_real, depth, balances in (("entry_13.csv", 13, 4, [385969, 459661]), ("entry_17.csv", 17, 5, [556863, 556863])):
        t = sb.MerkleSumTree.from_csv(os.path.join(golden_dir, name), ctx)
        o = M.MerkleSumTree.from_csv(os.path.join(golden_dir, name))
        root = t.root()
        assert t.depth() == depth and root.balances == balances and root.hash != 0
        assert (root.hash, root.balances) == (o.root[0], o.root[1])
        zero_leaf = M.poseidon_hash([0, 0, 0])
        for i in range(n_real, 1 << depth):
            assert t.node(0, i).hash == zero_leaf and t.node(0, i).balances == [0, 0]     # Entry::zero_entry
        proofs = [t.generate_proof(i) for i in range(1 << depth)]
        assert all(t.verify_proofs(proofs))
        with pytest.raises((SummaB200Error, IndexError, AssertionError)):
            t.generate_proof(1 << depth)
        t.close()


def test_biguint_balances_entry_16_bigints(ctx, golden_dir):
    """entry.rs:10 balances are BigUint; csv/entry_16_bigints.csv:2 holds 18446744073709551616 = 2^64 (merkle_sum_tree/tests.rs:130 `test_big_uint_conversion`).
    The wide entry point builds the same tree as the oracle; update_leaf accepts a wide balance too and restores the original root afterwards."""
    import circuits_halo2_b200 as sb
    path = os.path.join(golden_dir, "entry_16_bigints.csv")
    t = sb.MerkleSumTree.from_csv(path, ctx)
    o = M.MerkleSumTree.from_csv(path)
    assert sum(b >= (1 << 64) for e in o.entries for b in e.balances) >= 1
    root = t.root()
    assert (root.hash, root.balances) == (o.root[0], o.root[1])
    assert root.balances == [sum(e.balances[c] for e in o.entries) for c in range(2)] and max(root.balances) > (1 << 64)
    for lvl in range(o.depth + 1):
        assert [t.node(lvl, i).hash for i in range(len(o.nodes[lvl]))] == [h for h, _ in o.nodes[lvl]]
    assert all(t.verify_proofs([t.generate_proof(i) for i in range(16)]))
    big = (1 << 200) + 12345
    new_root = t.update_leaf(o.entries[3].username, [big, 7])
    ents = [M.Entry(e.username, list(e.balances)) for e in o.entries]
    ents[3] = M.Entry(ents[3].username, [big, 7])
    o2 = M.MerkleSumTree(ents)
    assert (new_root.hash, new_root.balances) == (o2.root[0], o2.root[1])
    back = t.update_leaf(o.entries[3].username, o.entries[3].balances)
    assert (back.hash, back.balances) == (o.root[0], o.root[1])
    t.close()
