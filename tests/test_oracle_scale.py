"""CPU tests of the threaded half of the oracle (oracle/halo2_cpu.c second half) and of the golden proofs it produced.

 * every C helper the large-k oracle prover uses is checked against its readable python twin (Poseidon / Merkle sum tree, Keccak-256,
   ChaCha20 `Fr::random`, `permute_expression_pair`, chunked Horner / batch inversion, powers, fixed-base multiplication);
 * the row-parallel evaluate_h program equals the term-by-term column-wise evaluation on the real circuit;
 * the committed golden proofs of BASELINE's K = 17 and K = 20 (tests/golden/make_golden_proofs.py) are ACCEPTED by the reference's
   verifier contract and tampered copies are rejected -- so the bytes the GPU must reproduce are valid proofs, not just oracle output;
 * the verifier fixture is the reference's file."""
import json
import os
import random

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import cpu
from oracle import halo2_prover as HP
from oracle import mst as M
from oracle import mst_circuit as C
from oracle import reference_verifier as RV
from oracle.chacha import ChaCha20Rng
from oracle.transcript import KeccakTranscript

um = lambda x: B.fr_from_mont_bytes(np.ascontiguousarray(x).tobytes())


def test_verifier_fixture_is_the_reference_file():
    ref = "/root/reference/contracts/src/InclusionVerifier.sol"
    if not os.path.exists(ref):
        pytest.skip("reference tree not mounted (GPU box)")
    assert open(ref, "rb").read() == open(RV.SOL, "rb").read()


def test_c_merkle_sum_tree_matches_python_twin():
    cpu.set_threads(4)
    for n in (1, 2, 5, 16, 33):
        ents = [(f"user_{i}", [100 + i, 7 * i, 3]) for i in range(n)]
        o = M.MerkleSumTree([M.Entry(u, b) for u, b in ents])
        t = cpu.MstC([u.encode() for u, _ in ents], np.array([b for _, b in ents], dtype=np.uint64))
        assert t.depth == o.depth
        for lvl in range(o.depth + 1):
            for i, (hh, bb) in enumerate(o.nodes[lvl]):
                h, b = t.node(lvl, i)
                assert um(h) == hh and [um(x) for x in b] == bb


def test_c_keccak_chacha_powers_match_python_twins():
    from oracle.keccak import keccak256
    for msg in (b"", b"dxGaEAii", b"x" * 135, b"y" * 136, b"z" * 300):
        assert cpu.keccak256(msg) == keccak256(msg)
    r = ChaCha20Rng(bytes(range(32)))
    _ = [r.next_fr() for _ in range(3)]
    exp = [r.next_fr() for _ in range(5)]
    assert [um(x) for x in cpu.chacha_fr_fill(ChaCha20Rng(bytes(range(32))).key, 3, 5)] == exp
    w = np.frombuffer(B.fr_to_mont_bytes(B.omega_for(12)), dtype=np.uint64)
    pw = cpu.fr_powers(w, 9001)
    assert [um(pw[i]) for i in (0, 1, 4095, 4096, 9000)] == [pow(B.omega_for(12), i, B.R) for i in (0, 1, 4095, 4096, 9000)]


def test_c_permute_expression_pair_matches_python_twin():
    rnd = random.Random(1)
    for usable in (1, 7, 1000):
        tab = (list(range(256)) + [0] * 1000)[:usable]
        inp = [rnd.randrange(min(256, usable)) for _ in range(usable)]
        pi, pt = HP.permute_expression_pair(inp, tab, usable)
        gi, gt = cpu.permute_expression_pair(HP.from_ints(inp), HP.from_ints(tab), usable)
        assert HP.to_ints(gi) == pi and HP.to_ints(gt) == pt
    with pytest.raises(ValueError):
        cpu.permute_expression_pair(HP.from_ints([300, 1, 2]), HP.from_ints([0, 1, 2]), 3)


def test_threaded_vector_helpers_match_serial():
    cpu.set_threads(4)
    a, b = cpu.random_fr(70001, 3), cpu.random_fr(70001, 4)
    a[::11] = 0
    x = cpu.random_fr(1, 5)[0]
    assert (cpu.par_fr_mul(a, b) == cpu.fr_mul(a, b)).all() and (cpu.par_fr_sub(a, b) == cpu.fr_sub(a, b)).all()
    assert (cpu.par_fr_eval_poly(a.reshape(-1), x) == cpu.fr_eval_poly(a.reshape(-1), x)).all()
    assert (cpu.par_fr_batch_invert(a.reshape(-1)) == cpu.fr_batch_invert(a.reshape(-1))).all()
    s = cpu.random_fr(50, 6)
    pts = cpu.g1_fixed_base_mul(s.reshape(-1))
    g = np.frombuffer(B.g1_to_mont_bytes((1, 2)), dtype=np.uint64)
    for i in (0, 17, 49):
        assert (pts[i] == cpu.g1_mul(g, s[i])).all()


def test_setup_srs_is_a_kzg_srs_and_program_h_equals_columnwise(golden_dir):
    """oracle `Params.setup` (tau known): commit(lagrange_to_coeff(v)) == commit_lagrange(v); and the two evaluate_h forms agree on the real
    circuit (same proof bytes)."""
    cpu.set_threads(4)
    k = 11
    params = HP.Params.setup(k, 0x5A110000 + k, threads=4)
    v = cpu.random_fr(1 << k, 9)
    dom = cpu.Domain(6, k, threads=4)
    assert params.commit(dom.lagrange_to_coeff(v.reshape(-1)).reshape(-1, 4)) == params.commit_lagrange(v)
    fx = np.load(os.path.join(golden_dir, "mst_inclusion_assignment.npz"))
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    pk = HP.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234)
    # the sparse constructor equals the dense one over the oracle's own synthesis
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(k, tree.generate_proof(0), 4, 2, 8)
    dense = HP.ProvingKey(params, cs, np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)]), C.permutation_mapping(lay), transcript_repr=0x1234)
    assert pk.fixed_commitments == dense.fixed_commitments and pk.sigma_commitments == dense.sigma_commitments
    adv = np.zeros((3, 1 << k, 4), dtype=np.uint64)
    adv[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
    inst = [um(x) for x in fx["instances"]]
    proofs = []
    for columnwise in (False, True):
        tr = KeccakTranscript()
        HP.create_proof(params, pk, inst, adv, ChaCha20Rng.seed_from_u64(1), tr, trace={"columnwise_h": columnwise})
        proofs.append(tr.finalize())
    assert proofs[0] == proofs[1]


@pytest.mark.parametrize("k", [17, 20])
def test_golden_proofs_are_accepted_by_the_reference_verifier(golden_dir, k):
    gold = np.load(os.path.join(golden_dir, f"golden_proof_k{k}.npz"))
    proof = gold["proof"].tobytes()
    args = (k, int(gold["tau"][0]), gold["fixed_comms"], gold["sigma_comms"], int(gold["transcript_repr"][0]))
    assert len(proof) == 2144
    assert RV.verify_mont(*args, proof, gold["instances"])
    bad = bytearray(proof)
    bad[0x400] ^= 1
    assert not RV.verify_mont(*args, bytes(bad), gold["instances"])
    inst = gold["instances"].copy()
    inst[2, 0] ^= np.uint64(1)
    assert not RV.verify_mont(*args, proof, inst)


def test_oracle_reproduces_entry_13_17_and_bigints_goldens(golden_dir):
    """merkle_sum_tree/tests.rs:219,251: root balances of csv/entry_13.csv (385969, 459661; depth 4) and csv/entry_17.csv (556863 twice; depth 5, the
    file ends with an empty line the csv crate skips); csv/entry_16_bigints.csv carries 2^64 (tests.rs:130)."""
    o13 = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_13.csv"))
    o17 = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_17.csv"))
    ob = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16_bigints.csv"))
    assert (o13.depth, o13.root[1]) == (4, [385969, 459661])
    assert (o17.depth, o17.root[1]) == (5, [556863, 556863])
    assert ob.depth == 4 and max(b for e in ob.entries for b in e.balances) == 1 << 64
    if os.path.exists("/root/reference/csv/entry_17.csv"):
        for f in ("entry_13.csv", "entry_17.csv", "entry_16_bigints.csv"):
            assert open(os.path.join(golden_dir, f), "rb").read() == open("/root/reference/csv/" + f, "rb").read()
