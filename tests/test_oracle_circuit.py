"""Pins the oracle's restatement of the reference circuit and prover (CPU only):

 * Merkle-sum-tree / Poseidon known answers quoted by the Rust tests;
 * keygen: the 11 fixed and 6 permutation commitments embedded in the reference's verifier contract
    are reproduced bit-exactly from csv/entry_16.csv + backend/ptau/hermez-raw-11, which pins
   the floor planner, the Pow5 / range / MST chip layouts, selector compression and Assembly::copy;
 * create_proof: a fresh oracle proof is ACCEPTED by the reference's own verifier
   (contracts/src/InclusionVerifier.sol run by oracle/yul.py), the checked-in golden proof
   is accepted, tampered proofs / instances are rejected.  The contract text is the committed fixture
   tests/golden/InclusionVerifier.sol (byte-identical to the reference's file: test_oracle_golden.py)."""
import json
import os

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import halo2_prover as HP
from oracle import mst as M
from oracle import mst_circuit as C
from oracle.chacha import ChaCha20Rng
from oracle.transcript import Blake2bTranscript, KeccakTranscript

from oracle.reference_verifier import SOL  # the reference contract, committed unchanged as a fixture

needs_reference = pytest.mark.skipif(not os.path.exists(SOL), reason="verifier fixture missing")


@pytest.fixture(scope="module")
def setup(golden_dir):
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(11, tree.generate_proof(0), 4, 2, 8)
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    params = HP.Params.read(os.path.join(golden_dir, "hermez-raw-11"))
    pk = HP.ProvingKey(params, cs, np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)]), C.permutation_mapping(lay))
    advice = np.stack([HP.from_ints(c) for c in C.advice_columns(lay)])
    instances = [tree.nodes[0][0][0], tree.root[0]] + tree.root[1]
    return dict(tree=tree, lay=lay, cs=cs, params=params, pk=pk, advice=advice, instances=instances)


def test_mst_known_answers(setup, golden_dir):
    gold = json.load(open(os.path.join(golden_dir, "mst_hashes.json")))
    t = setup["tree"]
    assert hex(t.nodes[0][0][0]) in gold["circuit_tests_hex"]      # circuits/tests.rs:341
    assert hex(t.nodes[0][1][0]) in gold["circuit_tests_hex"]      # circuits/tests.rs:346
    assert hex(t.root[0]) in gold["backend_tests_hex"]             # backend/src/tests.rs:265
    assert t.root[1] == [556862, 556862]                           # merkle_sum_tree/tests.rs:24


def test_layout_matches_mockprover_region_indices(setup):
    lay = setup["lay"]
    assert len(lay.region_starts) == 122          # last `permute state` region is #121 (circuits/tests.rs:113)
    assert lay.region_starts[21] == 246           # first range-check region (#21) / constant at fixed col 2 row 246 (tests.rs:287-288)
    assert max(lay.next_free.values()) <= (1 << 11) - 6


def test_keygen_reproduces_the_reference_verifying_key(setup, golden_dir):
    vk = json.load(open(os.path.join(golden_dir, "verifier_constants.json")))
    pk = setup["pk"]
    for i, c in enumerate(pk.fixed_commitments):
        assert c == (int(vk[f"fixed_comms[{i}].x"], 16), int(vk[f"fixed_comms[{i}].y"], 16)), f"fixed_comms[{i}]"
    for j, c in enumerate(pk.sigma_commitments):
        assert c == (int(vk[f"permutation_comms[{j}].x"], 16), int(vk[f"permutation_comms[{j}].y"], 16)), f"permutation_comms[{j}]"


def test_constraint_system_shape(setup):
    cs = setup["cs"]
    assert (cs["num_advice_columns"], cs["num_fixed_columns"], cs["num_instances"]) == (3, 11, 4)
    assert len(cs["gates"]) == 19 and len(cs["lookups"]) == 1
    assert cs["permutation_columns"] == [["fixed", 2], ["advice", 0], ["advice", 1], ["fixed", 3], ["advice", 2], ["instance", 0]]
    assert cs["advice_queries"] == [[0, 0], [1, 0], [0, 1], [1, 1], [2, 0], [1, -1], [0, -1]]


def test_witness_satisfies_every_gate_and_lookup(setup):
    """MockProver-style check on the Lagrange domain (usable rows)."""
    cs, pk, adv, n = setup["cs"], setup["pk"], setup["advice"], 1 << 11
    inst = np.zeros((n, 4), dtype=np.uint64)
    col = lambda kind, c, r: HP.rot(adv[c] if kind == "advice" else pk.fixed_values[c] if kind == "fixed" else inst, r)
    usable = n - 6
    for gi, g in enumerate(cs["gates"]):
        vals = HP.eval_expr(g, col, n)
        assert not vals[:usable - 1].any(), f"gate {gi} not satisfied"
    lk = cs["lookups"][0]
    inp = set(HP.to_ints(HP.eval_expr(lk["input"][0], col, n)[:usable]))
    tab = set(HP.to_ints(HP.eval_expr(lk["table"][0], col, n)[:usable]))
    assert inp <= tab


def test_chacha_rng_known_answers():
    assert ChaCha20Rng(bytes(32)).next_u32() == 0xADE0B876   # first keystream word of the all-zero key
    r = ChaCha20Rng.seed_from_u64(0)
    assert 0 <= r.next_fr() < B.R


@needs_reference
def test_reference_verifier_accepts_golden_and_oracle_proofs(setup, golden_dir):
    from oracle.yul import SolidityVerifier
    v = SolidityVerifier.from_file(SOL)
    cd = json.load(open(os.path.join(golden_dir, "inclusion_proof_solidity_calldata.json")))
    gproof = bytes.fromhex(cd["proof"][2:])
    ginst = [int(x, 16) for x in cd["public_inputs"]]
    assert v.verify(gproof, ginst)
    bad = bytearray(gproof)
    bad[0x380 + 5] ^= 1
    assert not v.verify(bytes(bad), ginst)
    # fresh oracle proof for the current circuit, real SRS, real vk
    tr = KeccakTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(42), tr)
    proof = tr.finalize()
    assert len(proof) == 2144
    assert v.verify(proof, setup["instances"])
    wrong = list(setup["instances"])
    wrong[2] += 1
    assert not v.verify(proof, wrong)
    # determinism in the RNG seed, sensitivity to it
    tr2 = KeccakTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(42), tr2)
    assert tr2.finalize() == proof
    tr3 = KeccakTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(43), tr3)
    assert tr3.finalize() != proof and v.verify(tr3.finalize(), setup["instances"])


def test_blake2b_transcript_proof_shape(setup):
    """full_prover's transcript (utils.rs:93): 16 compressed points + 35 scalars = 1632 bytes."""
    tr = Blake2bTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(1), tr)
    assert len(tr.finalize()) == 16 * 32 + 35 * 32


@needs_reference
def test_generic_verifier_agrees_with_the_reference_contract(setup, golden_dir):
    """oracle/halo2_verifier.py (any constraint system) vs the reference's verifier contract (2-currency circuit) on the REAL SRS / vk at k = 11,
    pairing path with [s]_2 from the contract's constants: the checked-in golden calldata proof, a fresh oracle proof, tampered proofs, wrong
    instances -- same verdict every time.  And a Blake2b-transcript proof (`full_prover`, utils.rs:93), which no contract can judge, verifies too."""
    from oracle import halo2_verifier as V
    from oracle.yul import SolidityVerifier
    v = SolidityVerifier.from_file(SOL)
    vc = json.load(open(os.path.join(golden_dir, "verifier_constants.json")))
    g = lambda name: int(vc[name], 16)
    s_g2 = ((g("neg_s_g2_x_2"), g("neg_s_g2_x_1")), ((-g("neg_s_g2_y_2")) % B.Q, (-g("neg_s_g2_y_1")) % B.Q))   # the contract stores -[s]_2, imaginary part first
    fixed = [(g(f"fixed_comms[{i}].x"), g(f"fixed_comms[{i}].y")) for i in range(11)]
    sigma = [(g(f"permutation_comms[{i}].x"), g(f"permutation_comms[{i}].y")) for i in range(6)]
    cs = setup["cs"]
    check = lambda proof, inst: V.verify_proof(cs, 11, fixed, sigma, g("vk_digest"), inst, proof, keccak=True, s_g2=s_g2)
    cd = json.load(open(os.path.join(golden_dir, "inclusion_proof_solidity_calldata.json")))
    gproof, ginst = bytes.fromhex(cd["proof"][2:]), [int(x, 16) for x in cd["public_inputs"]]
    tr = KeccakTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(7), tr)
    fresh = tr.finalize()
    cases = [(gproof, ginst), (fresh, setup["instances"]), (fresh, ginst)]
    for off in (0x10, 0x150, 0x250, 0x390, 0x700, 0x7f0, 0x830):
        bad = bytearray(fresh)
        bad[off] ^= 1
        cases.append((bytes(bad), setup["instances"]))
    cases.append((fresh[:-1], setup["instances"]))
    verdicts = [(check(p, i), v.verify(p, i) if len(p) == 2144 else False) for p, i in cases]
    assert all(a == b for a, b in verdicts), verdicts
    assert verdicts[0] == (True, True) and verdicts[1] == (True, True) and not any(a for a, _ in verdicts[2:])
    # the Blake2b transcript (compressed points, little-endian scalars)
    trb = Blake2bTranscript()
    HP.create_proof(setup["params"], setup["pk"], setup["instances"], setup["advice"], ChaCha20Rng.seed_from_u64(3), trb)
    bproof = trb.finalize()
    okb = lambda p, inst: V.verify_proof(cs, 11, fixed, sigma, setup["pk"].transcript_repr, inst, p, keccak=False, s_g2=s_g2)
    assert okb(bproof, setup["instances"])
    bad = bytearray(bproof)
    bad[600] ^= 1
    assert not okb(bytes(bad), setup["instances"])
