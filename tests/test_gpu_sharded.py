"""GPU tests of the sharded create_proof: coset-sharded evaluate_h + base-range-sharded MSMs.

world = 1 exercises the coset path on one GPU against the oracle prover; world = 2 (when the box has two GPUs) runs one
process per GPU under torchrun and requires byte equality with the single-GPU proof."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle import halo2_prover as HP
from oracle import mst as M
from oracle import mst_circuit as C
from oracle.chacha import ChaCha20Rng
from oracle.transcript import KeccakTranscript

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_coset_path_world1_is_byte_identical_to_oracle(ctx, golden_dir):
    import circuits_halo2_b200 as sb
    tree = M.MerkleSumTree.from_csv(os.path.join(golden_dir, "entry_16.csv"))
    lay = C.synthesize(11, tree.generate_proof(5), 4, 2, 8)
    cs = json.load(open(os.path.join(golden_dir, "mst_inclusion_cs.json")))
    oparams = HP.Params.read(os.path.join(golden_dir, "hermez-raw-11"))
    fixed = np.stack([HP.from_ints(c) for c in C.fixed_columns(lay)])
    opk = HP.ProvingKey(oparams, cs, fixed, C.permutation_mapping(lay))
    advice = np.stack([HP.from_ints(c) for c in C.advice_columns(lay)])
    instances = [tree.nodes[0][5][0], tree.root[0]] + tree.root[1]
    params = sb.ParamsKZG.read(os.path.join(golden_dir, "hermez-raw-11"), ctx)
    pk = sb.ProvingKey(params, cs, fixed, opk.sigma_values, opk.transcript_repr, ctx)
    tr = KeccakTranscript()
    HP.create_proof(oparams, opk, instances, advice, ChaCha20Rng.seed_from_u64(11), tr)
    got = sb.create_proof(pk, instances, advice, sb.seed_from_u64(11), sb.TRANSCRIPT_KECCAK, comm=sb.LocalComm())
    assert got == tr.finalize()


@pytest.mark.parametrize("k", [13, 16])
def test_coset_path_world1_equals_plain_path(ctx, golden_dir, k):
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import fields
    fx = np.load(os.path.join(golden_dir, "mst_inclusion_assignment.npz"))
    cs = open(os.path.join(golden_dir, "mst_inclusion_cs.json")).read()
    params = sb.ParamsKZG.setup(k, 0x5A110000 + k, ctx, download=False)
    pk = sb.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234, ctx)
    advice = np.zeros((3, 1 << k, 4), dtype=np.uint64)
    advice[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
    instances = [fields.fr_from_mont(v) for v in fx["instances"]]
    for tk in (sb.TRANSCRIPT_KECCAK, sb.TRANSCRIPT_BLAKE2B):
        assert sb.create_proof(pk, instances, advice, sb.seed_from_u64(k), tk, comm=sb.LocalComm()) == sb.create_proof(pk, instances, advice, sb.seed_from_u64(k), tk)


@pytest.mark.parametrize("log_n", [16, 17, 20, 22])
def test_distributed_ntt_world1_equals_local(ctx, log_n):
    """sb_ntt_dist with a single rank (no exchange): the tile-range launches of both passes reproduce best_fft, with and without the folded scale"""
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import fields
    a = np.random.default_rng(log_n).integers(0, 1 << 62, size=(1 << log_n, 4), dtype=np.uint64)
    a[:, 3] &= np.uint64((1 << 61) - 1)   # < 2^253 < r: canonical residues
    w = fields.fr_to_mont(fields.omega(log_n))
    want = sb.best_fft(a.copy(), w, log_n, ctx)
    assert (sb.best_fft_dist(a.copy(), w, log_n, sb.LocalComm(), ctx) == want).all()
    d = sb.EvaluationDomain(3, log_n, ctx)
    winv = fields.fr_to_mont(pow(fields.omega(log_n), -1, fields.FR_MODULUS))
    ninv = fields.fr_to_mont(pow(1 << log_n, -1, fields.FR_MODULUS))
    assert (sb.best_fft_dist(want.copy(), winv, log_n, sb.LocalComm(), ctx, scale=ninv) == a).all()          # inverse with n^-1 folded in: round trip
    assert (d.lagrange_to_coeff(want.copy()) == a).all()


def test_sharded_rejects_comm_without_callbacks(ctx, golden_dir):
    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import fields
    from circuits_halo2_b200._lib import SummaB200Error
    fx = np.load(os.path.join(golden_dir, "mst_inclusion_assignment.npz"))
    cs = open(os.path.join(golden_dir, "mst_inclusion_cs.json")).read()
    params = sb.ParamsKZG.setup(12, 1, ctx, download=False)
    pk = sb.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 1, ctx)
    advice = np.zeros((3, 1 << 12, 4), dtype=np.uint64)
    bad = sb.LocalComm()
    bad.struct.world, bad.struct.rank = 3, 0
    with pytest.raises(SummaB200Error):
        sb.create_proof(pk, [fields.fr_from_mont(v) for v in fx["instances"]], advice, sb.seed_from_u64(1), sb.TRANSCRIPT_KECCAK, comm=bad)


def test_two_gpu_sharded_proof_equals_single_gpu_proof():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1", "--master-port", "29533",
           os.path.join(ROOT, "tests", "multi", "sharded_proof_worker.py"), "14"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    # the library's own communicator (shared-memory mailbox for host records, CUDA IPC peer copies for device buffers) instead of torch / NCCL,
    # including the distributed NTT's all-to-all
    cmd[cmd.index("29533")] = "29536"
    out = subprocess.run(cmd[:-1] + ["15", "1", "shm"], capture_output=True, text=True, timeout=600, env=dict(os.environ, SB_DIST_NTT_MIN_K="10"))
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    cmd[cmd.index("29536")] = "29533"
    # the same with every replicated size-n transform run as a distributed four-step NTT (the k >= 22 path, forced at k = 16)
    env = dict(os.environ, SB_DIST_NTT_MIN_K="10")
    cmd[cmd.index("29533")] = "29535"
    out = subprocess.run(cmd[:-1] + ["16"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
    cmd[cmd.index("29535")] = "29533"
    # the north_star's other split of a commitment: by base range, 64-byte partial points added on the host
    env = dict(os.environ, SB_SHARD_MSM_BY_RANGE="1")
    cmd[cmd.index("29533")] = "29534"
    out = subprocess.run(cmd[:-1] + ["13"], capture_output=True, text=True, timeout=600, env=env)
    assert out.returncode == 0 and "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-4000:]
