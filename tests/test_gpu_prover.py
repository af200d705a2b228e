"""GPU parity tests of the create_proof pipeline (python -m pytest tests -m gpu).

Rows a6-a9 of SURVEY 8(a): batch inversion, grand-product scan, lookup permutation (sort), polynomial
evaluation, kate division, then the whole `create_proof`: the GPU prover's proof must be BYTE-IDENTICAL to
the oracle prover's (oracle/halo2_prover.py, itself accepted by the reference's verifier contract) for the
same circuit, SRS and ChaCha20 seed, under both transcripts.  Bit-exact; no tolerance."""
import ctypes
import json
import os

import numpy as np
import pytest

from oracle import bn254 as B
from oracle import cpu
from oracle import halo2_prover as HP
from oracle import mst as M
from oracle import mst_circuit as C
from oracle.chacha import ChaCha20Rng
from oracle.transcript import Blake2bTranscript, KeccakTranscript

pytestmark = pytest.mark.gpu


def fr(x):
    return np.frombuffer(B.fr_to_mont_bytes(x), dtype=np.uint64).copy()


def L():
    from circuits_halo2_b200 import _lib
    return _lib


def P(a):
    from circuits_halo2_b200.context import ptr
    return ptr(a)


@pytest.mark.parametrize("n", [1, 31, 32, 33, 1000, 1 << 16, (1 << 16) + 77, 300001])  # >= 2^16: the two-level kernels
def test_batch_invert(ctx, n):
    a = cpu.random_fr(n, 11 + n)
    a[::7] = 0  # zeros stay zero (halo2 batch_invert)
    got = a.copy()
    L().check(L().lib().sb_fr_batch_invert(ctx.handle, P(got), ctypes.c_size_t(n)), "batch_invert")
    assert (got.reshape(-1) == cpu.fr_batch_invert(a)).all()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 65, 5000, (1 << 17) + 3])
def test_running_product(ctx, n):
    a = cpu.random_fr(n, 21 + n)
    init = fr(12345)
    z = np.zeros((n, 4), dtype=np.uint64)
    L().check(L().lib().sb_fr_running_product(ctx.handle, P(a), ctypes.c_size_t(n), P(init), P(z), ctypes.c_size_t(n)), "running_product")
    assert (z.reshape(-1) == cpu.fr_running_product(a, init, n)).all()


@pytest.mark.parametrize("n", [1, 255, 256, 257, 1 << 12, (1 << 16) + 17])
def test_eval_polynomial(ctx, n):
    coeffs = cpu.random_fr(n, 31 + n)
    xs = [0, 1, 7, B.R - 1, B.omega_for(10), 0x1234567890ABCDEF1234567890ABCDEF]
    pts = np.concatenate([fr(x) for x in xs]).reshape(-1, 4)
    out = np.zeros((len(xs), 4), dtype=np.uint64)
    L().check(L().lib().sb_fr_eval_polynomial(ctx.handle, P(coeffs), ctypes.c_size_t(n), P(pts), ctypes.c_size_t(len(xs)), P(out)), "eval")
    for i in range(len(xs)):
        assert (out[i] == cpu.fr_eval_poly(coeffs, pts[i])).all(), (n, i)


@pytest.mark.parametrize("n", [1, 2, 3, 100, 1024, 1025, 5000, 1 << 15])
def test_sort_matches_fr_ord(ctx, n):
    a = cpu.random_fr(n, 41 + n)
    a[: n // 3] = a[0]                      # duplicates
    if n > 10:
        a[5] = fr(0); a[6] = fr(B.R - 1); a[7] = fr(1)
    got = a.copy()
    L().check(L().lib().sb_fr_sort(ctx.handle, P(got), ctypes.c_size_t(n)), "sort")
    ref = sorted(HP.to_ints(a))
    assert HP.to_ints(got) == ref


@pytest.mark.parametrize("n,top", [(4096, 255), (5000, 65535), (1 << 16, 65535), (5000, 65536), ((1 << 15) + 7, 1 << 200)])
def test_sort_small_keys_counting_path_and_fallback(ctx, n, top):
    """keys below 2^16 take the histogram path (byte tables, range-check inputs); a single larger key falls back to the bitonic network"""
    rng = np.random.default_rng(n + top % 1000)
    vals = [int(x) for x in rng.integers(0, 256, size=n)]
    vals[0], vals[1], vals[n // 2] = 0, top, top
    vals[n // 3: n // 3 + 700] = [7] * 700      # a hot bin
    a = np.stack([fr(v) for v in vals])
    got = a.copy()
    L().check(L().lib().sb_fr_sort(ctx.handle, P(got), ctypes.c_size_t(n)), "sort")
    assert HP.to_ints(got) == sorted(vals)


@pytest.mark.parametrize("n,kind", [(256, "bytes"), (2048, "bytes"), (2048, "random"), (1 << 14, "bytes")])
def test_lookup_permute(ctx, n, kind):
    rng = np.random.default_rng(n)
    usable = n - 6
    table_vals = list(range(256)) + [0] * (n - 256) if kind == "bytes" else [int(x) for x in rng.integers(1, 1 << 62, size=n)]
    if kind == "bytes":
        inp_vals = [int(x) for x in rng.integers(0, min(256, usable), size=n)]  # only values present in the usable table rows
        inp_vals[: n // 2] = [0] * (n // 2)
    else:
        inp_vals = [table_vals[int(i)] for i in rng.integers(0, usable, size=n)]
    inp, tab = HP.from_ints(inp_vals), HP.from_ints(table_vals)
    a_ref, s_ref = HP.permute_expression_pair(inp_vals, table_vals, usable)
    a_out = np.zeros((usable, 4), dtype=np.uint64)
    s_out = np.zeros((usable, 4), dtype=np.uint64)
    L().check(L().lib().sb_lookup_permute(ctx.handle, P(inp), P(tab), ctypes.c_size_t(n), ctypes.c_size_t(usable), P(a_out), P(s_out)), "lookup_permute")
    assert HP.to_ints(a_out) == a_ref
    assert HP.to_ints(s_out) == s_ref


def test_lookup_permute_rejects_value_outside_table(ctx):
    from circuits_halo2_b200._lib import SummaB200Error
    n, usable = 256, 250
    inp = HP.from_ints([999] + [1] * (n - 1))
    tab = HP.from_ints(list(range(n)))
    out = np.zeros((usable, 4), dtype=np.uint64)
    with pytest.raises(SummaB200Error):
        L().check(L().lib().sb_lookup_permute(ctx.handle, P(inp), P(tab), ctypes.c_size_t(n), ctypes.c_size_t(usable), P(out), P(out.copy())), "lookup_permute")


@pytest.mark.parametrize("log_n", [7, 11, 16])
def test_kate_division(ctx, log_n):
    n = 1 << log_n
    a = cpu.random_fr(n, 51 + log_n)
    b = fr(0xDEADBEEF1234567)
    q = np.zeros((n, 4), dtype=np.uint64)
    L().check(L().lib().sb_kate_division(ctx.handle, P(a), ctypes.c_uint32(log_n), P(b), P(q)), "kate")
    ref = cpu.fr_kate_division(a, b).reshape(-1, 4)
    assert (q[: n - 1] == ref).all() and not q[n - 1].any()


# ------------------------------------------------------------------ the real circuit, real SRS (k = 11)
@pytest.fixture(scope="module")
def real(golden_dir, ctx):
    import circuits_halo2_b200 as sb
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(11, tree.generate_proof(0), 4, 2, 8)
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    oparams = HP.Params.read(os.path.join(golden_dir, "hermez-raw-11"))
    fixed = np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)])
    opk = HP.ProvingKey(oparams, cs, fixed, C.permutation_mapping(lay))
    advice = np.stack([HP.from_ints(c) for c in C.advice_columns(lay)])
    instances = [tree.nodes[0][0][0], tree.root[0]] + tree.root[1]
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    pk = sb.ProvingKey(params, cs, fixed, opk.sigma_values, opk.transcript_repr, ctx)
    return dict(cs=cs, oparams=oparams, opk=opk, advice=advice, instances=instances, params=params, pk=pk)


def test_pk_commitments_equal_reference_vk(real, golden_dir):
    """keygen commitments computed on the GPU == the constants of the reference's verifier contract (.sol:238-271)."""
    vk = json.load(open(os.path.join(golden_dir, "verifier_constants.json")))
    f, s = real["pk"].commitments()
    for i in range(11):
        assert B.g1_from_mont_bytes(f[i].tobytes()) == (int(vk[f"fixed_comms[{i}].x"], 16), int(vk[f"fixed_comms[{i}].y"], 16))
    for j in range(6):
        assert B.g1_from_mont_bytes(s[j].tobytes()) == (int(vk[f"permutation_comms[{j}].x"], 16), int(vk[f"permutation_comms[{j}].y"], 16))


@pytest.mark.parametrize("seed", [42, 7])
def test_create_proof_keccak_is_byte_identical_to_oracle(real, seed):
    import circuits_halo2_b200 as sb
    tr = KeccakTranscript()
    HP.create_proof(real["oparams"], real["opk"], real["instances"], real["advice"], ChaCha20Rng.seed_from_u64(seed), tr)
    ref = tr.finalize()
    got = sb.create_proof(real["pk"], real["instances"], real["advice"], sb.seed_from_u64(seed), sb.TRANSCRIPT_KECCAK)
    assert len(got) == 2144
    if got != ref:
        first = next(i for i in range(min(len(got), len(ref))) if got[i] != ref[i])
        pytest.fail(f"proof differs from the oracle's at byte {first:#x} (section boundaries: advice 0x0, lookup 0xc0, perm 0x140, lookupZ 0x1c0, "
                    f"random 0x200, h 0x240, evals 0x380, W 0x7e0, W' 0x820)")


def test_create_proof_blake2b_is_byte_identical_to_oracle(real):
    import circuits_halo2_b200 as sb
    tr = Blake2bTranscript()
    HP.create_proof(real["oparams"], real["opk"], real["instances"], real["advice"], ChaCha20Rng.seed_from_u64(3), tr)
    got = sb.create_proof(real["pk"], real["instances"], real["advice"], sb.seed_from_u64(3), sb.TRANSCRIPT_BLAKE2B)
    assert got == tr.finalize() and len(got) == 51 * 32


def test_create_proof_rejects_unsatisfiable_lookup(real):
    """a balance that does not fit N_BYTES makes a range-check input fall outside the table: the prover errors like halo2 does."""
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200._lib import SummaB200Error
    bad = real["advice"].copy()
    lay_row = 246  # first range-check region: z_0 at advice column 0 (circuits/tests.rs:287)
    bad[0, lay_row] = fr(B.R - 5)
    with pytest.raises(SummaB200Error):
        sb.create_proof(real["pk"], real["instances"], bad, sb.seed_from_u64(1), sb.TRANSCRIPT_KECCAK)


@pytest.mark.parametrize("k", [12, 13])
def test_create_proof_larger_k_synthetic_srs(ctx, golden_dir, k):
    """Same circuit in a larger domain, unsafe synthetic SRS generated on the GPU: byte-identical to the oracle."""
    import circuits_halo2_b200 as sb
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(k, tree.generate_proof(3), 4, 2, 8)
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    params = sb.ParamsKZG.setup(k, 0x5A110000 + k, ctx)
    oparams = HP.Params(k, params.g, params.g_lagrange, threads=8)
    # the synthetic SRS must be a KZG SRS: commit(lagrange_to_coeff(v)) == commit_lagrange(v)
    v = cpu.random_fr(1 << k, 5)
    dom = sb.EvaluationDomain(6, k, ctx)
    assert (params.commit(dom.lagrange_to_coeff(v)) == params.commit_lagrange(v)).all()
    fixed = np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)])
    opk = HP.ProvingKey(oparams, cs, fixed, C.permutation_mapping(lay), transcript_repr=0x1234)
    advice = np.stack([HP.from_ints(c) for c in C.advice_columns(lay)])
    instances = [tree.nodes[0][3][0], tree.root[0]] + tree.root[1]
    pk = sb.ProvingKey(params, cs, fixed, opk.sigma_values, 0x1234, ctx)
    tr = KeccakTranscript()
    HP.create_proof(oparams, opk, instances, advice, ChaCha20Rng.seed_from_u64(k), tr)
    got = sb.create_proof(pk, instances, advice, sb.seed_from_u64(k), sb.TRANSCRIPT_KECCAK)
    assert got == tr.finalize()


def test_sparse_key_and_device_srs_match_dense_path(ctx, golden_dir):
    """ProvingKey.from_sparse (fixture of tests/golden/make_assignment.py) + ParamsKZG.setup on the device give the same
    proof as the dense path, and that proof equals the oracle's at k = 14."""
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import fields
    k = 14
    n = 1 << k
    fx = np.load(os.path.join(golden_dir, "mst_inclusion_assignment.npz"))
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    params = sb.ParamsKZG.setup(k, 0xABCDEF, ctx)
    pk = sb.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x77, ctx)
    advice = np.zeros((3, n, 4), dtype=np.uint64)
    advice[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
    instances = [fields.fr_from_mont(v) for v in fx["instances"]]
    got = sb.create_proof(pk, instances, advice, sb.seed_from_u64(5), sb.TRANSCRIPT_KECCAK)
    # oracle over the same SRS bytes, circuit re-synthesised by the oracle at this k
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(k, tree.generate_proof(0), 4, 2, 8)
    oparams = HP.Params(k, params.g, params.g_lagrange, threads=8)
    fixed = np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)])
    opk = HP.ProvingKey(oparams, cs, fixed, C.permutation_mapping(lay), transcript_repr=0x77)
    f, s = pk.commitments()
    assert [B.g1_from_mont_bytes(x.tobytes()) for x in f] == opk.fixed_commitments
    assert [B.g1_from_mont_bytes(x.tobytes()) for x in s] == opk.sigma_commitments
    tr = KeccakTranscript()
    HP.create_proof(oparams, opk, instances, np.stack([HP.from_ints(c) for c in C.advice_columns(lay)]), ChaCha20Rng.seed_from_u64(5), tr)
    assert got == tr.finalize()


@pytest.mark.parametrize("user", [1, 9])
def test_end_to_end_gpu_tree_to_gpu_proof_for_other_users(real, golden_dir, ctx, user):
    """The whole product flow for users other than the fixture's: Merkle sum tree on the GPU -> its Merkle proof for `user` -> the circuit's
    witness (the Rust front-end's job; restated by oracle/mst_circuit.py) -> create_proof on the GPU with the key built for user 0.
    The proof must equal the oracle prover's byte for byte (that prover's proofs are accepted by the reference's verifier contract,
    tests/test_oracle_circuit.py) and carry the user's leaf hash and the tree's root as public inputs."""
    import circuits_halo2_b200 as sb

    class _E:  # what oracle/mst_circuit.py reads from the entry: its hash preimage
        def __init__(self, pre):
            self._pre = pre

        def preimage(self):
            return self._pre

    tree = sb.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"), ctx)
    mp = tree.generate_proof(user)
    assert tree.verify_proof(mp)
    proof_dict = {"entry": _E(mp.entry_preimage), "root": (mp.root.hash, mp.root.balances),
                  "sibling_leaf_node_hash_preimage": mp.sibling_leaf_node_hash_preimage,
                  "sibling_middle_node_hash_preimages": mp.sibling_middle_node_hash_preimages, "path_indices": mp.path_indices}
    lay = C.synthesize(11, proof_dict, 4, 2, 8)
    # the key does not depend on the user: fixed columns and permutation are those of the fixture's key
    assert (np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)]) == np.stack([HP.from_ints(c) for c in C.fixed_columns(C.synthesize(11, M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv")).generate_proof(0), 4, 2, 8))])).all()
    advice = np.stack([HP.from_ints(c) for c in C.advice_columns(lay)])
    instances = [tree.node(0, user).hash, mp.root.hash] + list(mp.root.balances)
    tr = KeccakTranscript()
    HP.create_proof(real["oparams"], real["opk"], instances, advice, ChaCha20Rng.seed_from_u64(100 + user), tr)
    got = sb.create_proof(real["pk"], instances, advice, sb.seed_from_u64(100 + user), sb.TRANSCRIPT_KECCAK)
    assert got == tr.finalize()


def test_batch_prover_equals_sequential_proofs(real):
    """BatchProver (one context per worker thread, one shared key): every proof equals the one the key's own context makes for the same
    seed, under both transcripts, and distinct seeds give distinct proofs."""
    import circuits_halo2_b200 as sb
    jobs = [(real["instances"], real["advice"], sb.seed_from_u64(500 + j), sb.TRANSCRIPT_KECCAK if j % 2 == 0 else sb.TRANSCRIPT_BLAKE2B) for j in range(12)]
    bp = sb.BatchProver(real["pk"], workers=4)
    try:
        got = bp.prove_many(jobs)
    finally:
        bp.close()
    ref = [sb.create_proof(real["pk"], *job) for job in jobs]
    assert got == ref and len(set(got)) == len(got)


def test_new_entry_points_reject_bad_arguments(ctx, golden_dir):
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import _lib
    from circuits_halo2_b200._lib import SummaB200Error
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    with pytest.raises(SummaB200Error):
        params.precompute(bases=0)                      # empty basis mask
    with pytest.raises(SummaB200Error):
        params.precompute(bases=1, window_bits=30)      # window wider than the tables support
    small = params.downsize(8)
    small.precompute()                                  # k < 11: a no-op, commits keep working
    v = cpu.random_fr(256, 3)
    assert (small.commit(v) == params.commit(v)).all()  # the first 2^8 monomial bases are shared
    assert not params.commit(np.zeros((0, 4), dtype=np.uint64)).any()   # empty polynomial -> identity, with or without tables
    params.precompute()
    assert not params.commit(np.zeros((0, 4), dtype=np.uint64)).any()
    with pytest.raises(AssertionError):
        sb.MerkleSumTree.from_entries([], ctx=ctx)
    with pytest.raises(SummaB200Error):
        sb.MerkleSumTree.from_arrays([b"a"], np.zeros((1, 33), dtype=np.uint64), ctx=ctx)   # more than 32 currencies
    t = sb.MerkleSumTree.from_entries([sb.Entry("solo", [5])], ctx=ctx)                    # a single entry: depth 0, the leaf is the root
    assert t.depth() == 0 and t.root().balances == [5] and t.root().hash == M.poseidon_hash(M.Entry("solo", [5]).preimage())
    p = t.generate_proof(0)
    assert p.path_indices == [] and p.sibling_middle_node_hash_preimages == []


def test_levels_20_circuit_fixture_proof_equals_oracle(ctx, golden_dir):
    """BASELINE configs[4]'s circuit, MstInclusionCircuit<20, 2, 8> at its minimum k = 13 (fixture of tests/golden/make_assignment_l20.py: one
    fabricated Merkle path of a 2^20-leaf tree): key from the sparse fixture, proof byte-identical to the oracle prover's over the same SRS."""
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import fields
    k, n = 13, 1 << 13
    fx = np.load(os.path.join(golden_dir, "mst_inclusion_assignment_l20.npz"))
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    assert int(fx["rows_used"][0]) <= n - 6
    params = sb.ParamsKZG.setup(k, 0x5A110000 + k, ctx)
    pk = sb.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234, ctx)
    advice = np.zeros((3, n, 4), dtype=np.uint64)
    advice[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
    instances = [fields.fr_from_mont(v) for v in fx["instances"]]
    got = sb.create_proof(pk, instances, advice, sb.seed_from_u64(20), sb.TRANSCRIPT_KECCAK)
    # the oracle's key from the same sparse data (dense columns rebuilt on the host)
    fixed = np.zeros((cs["num_fixed_columns"], n, 4), dtype=np.uint64)
    fixed[fx["fixed_cells"][:, 0], fx["fixed_cells"][:, 1]] = fx["fixed_values"]
    ncols = len(cs["permutation_columns"])
    mapping = [[(c, r) for r in range(n)] for c in range(ncols)]
    for c, r, tc, tr in fx["perm_cells"]:
        mapping[int(c)][int(r)] = (int(tc), int(tr))
    oparams = HP.Params(k, params.g, params.g_lagrange, threads=8)
    opk = HP.ProvingKey(oparams, cs, fixed, mapping, transcript_repr=0x1234)
    tr = KeccakTranscript()
    HP.create_proof(oparams, opk, instances, advice, ChaCha20Rng.seed_from_u64(20), tr)
    assert got == tr.finalize()


# ------------------------------------------------------------------ standalone boundary entries
@pytest.mark.parametrize("rot_scale_log,log_rows", [(3, 9), (0, 7)])
def test_evaluate_h_standalone_matches_oracle_numerator(ctx, golden_dir, rot_scale_log, log_rows):
    """sb_evaluate_h over RANDOM columns of the reference circuit's constraint system == the oracle's quotient numerator (oracle/halo2_prover.py
    `_h_numerator_program` / column-wise form), bit for bit: kernel-level parity of `Evaluator::evaluate_h`, both for halo2's extended-domain layout
    (rotation = 8 entries) and for a single coset (rotation = 1 entry)."""
    import types
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    rows = 1 << log_rows
    A, F, Pn = cs["num_advice_columns"], cs["num_fixed_columns"], len(cs["permutation_columns"])
    chunk = cs["degree"] - 2
    n_sets = -(-Pn // chunk)
    n_cols = A + F + 1 + Pn + n_sets + 4 + 3 * len(cs["lookups"])
    cols = [cpu.random_fr(rows, 900 + i) for i in range(n_cols)]
    theta, beta, gamma, y = [B.fr_from_mont_bytes(cpu.random_fr(1, 950 + i).tobytes()) for i in range(4)]
    E_SIGMA = A + F + 1
    E_PZ = E_SIGMA + Pn
    E_L0 = E_PZ + n_sets
    E_LK = E_L0 + 4
    pk = types.SimpleNamespace(fixed_cosets=cols[A:A + F], sigma_cosets=cols[E_SIGMA:E_SIGMA + Pn], l0=cols[E_L0], l_last=cols[E_L0 + 1], l_active_row=cols[E_L0 + 2],
                               x_coset=cols[E_L0 + 3])
    pcols = [tuple(c) for c in cs["permutation_columns"]]
    perm_sets = [dict(coset=cols[E_PZ + s], cols=pcols[s * chunk:(s + 1) * chunk], first=s * chunk) for s in range(n_sets)]
    lk_cosets = [[cols[E_LK + 3 * li + j] for j in range(3)] for li in range(len(cs["lookups"]))]
    want = HP._h_numerator_program(cs, pk, cols[:A], cols[A + F], perm_sets, lk_cosets, theta, beta, gamma, y, rows, 1 << rot_scale_log, cs["blinding_factors"])
    out = np.zeros((rows, 4), dtype=np.uint64)
    ptrs = (ctypes.c_void_p * n_cols)(*[c.ctypes.data for c in cols])
    text = json.dumps(cs).encode()
    L().check(L().lib().sb_evaluate_h(ctx.handle, ctypes.c_char_p(text), ptrs, ctypes.c_size_t(n_cols), ctypes.c_uint32(log_rows), ctypes.c_uint32(rot_scale_log),
                                      P(fr(theta)), P(fr(beta)), P(fr(gamma)), P(fr(y)), P(out)), "sb_evaluate_h")
    assert (out == want).all()
    # a wrong column count is an argument error, not a crash
    from circuits_halo2_b200._lib import SummaB200Error
    with pytest.raises(SummaB200Error):
        L().check(L().lib().sb_evaluate_h(ctx.handle, ctypes.c_char_p(text), ptrs, ctypes.c_size_t(n_cols - 1), ctypes.c_uint32(log_rows), ctypes.c_uint32(rot_scale_log),
                                          P(fr(theta)), P(fr(beta)), P(fr(gamma)), P(fr(y)), P(out)), "sb_evaluate_h")


def test_msm_batch_and_grand_product_entries(ctx, golden_dir):
    import circuits_halo2_b200 as sb
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    n, m = 1 << 11, 11
    sc = cpu.random_fr(n * m, 321).reshape(m, n, 4)
    sc[3, 100:] = 0          # a sparse column
    sc[5] = sc[5, 0]         # a constant column (one hot bucket per window)
    for precompute in (False, True):
        if precompute:
            params.precompute()
        for basis in (0, 1):
            out = np.zeros((m, 8), dtype=np.uint64)
            L().check(L().lib().sb_msm_g1_batch(ctx.handle, params.handle, ctypes.c_int32(basis), P(sc), ctypes.c_size_t(n), ctypes.c_size_t(m), P(out)), "sb_msm_g1_batch")
            bases = params.g if basis == 0 else params.g_lagrange
            for j in range(m):
                assert (out[j] == cpu.best_multiexp(sc[j], bases, threads=8)).all(), (precompute, basis, j)
    # grand product: z[i + 1] = z[i] * num[i] / den[i]
    for nn in (1, 77, 5000):
        num, den, init = cpu.random_fr(nn, 11), cpu.random_fr(nn, 12), fr(777)
        z = np.zeros((nn + 1, 4), dtype=np.uint64)
        L().check(L().lib().sb_grand_product(ctx.handle, P(num), P(den), ctypes.c_size_t(nn), P(init), P(z), ctypes.c_size_t(nn + 1)), "sb_grand_product")
        ratio = cpu.fr_mul(num.reshape(-1), cpu.fr_batch_invert(den.reshape(-1)))
        assert (z.reshape(-1) == cpu.fr_running_product(ratio, init, nn + 1)).all()
