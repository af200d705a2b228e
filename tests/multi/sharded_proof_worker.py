"""torchrun worker: the same create_proof on every rank, sharded over the ranks' GPUs (one process per GPU, NCCL).

Launched by tests/test_gpu_sharded.py and by hand:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi/sharded_proof_worker.py 14
Every rank checks that its sharded proof equals the unsharded proof it computes on its own GPU, byte for byte."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import circuits_halo2_b200 as sb  # noqa: E402
from circuits_halo2_b200 import fields  # noqa: E402


def main():
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 14
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = sb.Context(local)
    fx = np.load(os.path.join(ROOT, "tests", "golden", "mst_inclusion_assignment.npz"))
    cs = open(os.path.join(ROOT, "tests", "golden", "mst_inclusion_cs.json")).read()
    params = sb.ParamsKZG.setup(k, 0x5A110000 + k, ctx, download=False)
    pk = sb.ProvingKey.from_sparse(params, cs, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234, ctx)
    n = 1 << k
    advice = torch.zeros((3, n, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
    advice[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
    instances = [fields.fr_from_mont(v) for v in fx["instances"]]
    seed = sb.seed_from_u64(99)
    # argv[3] = "shm": the library's own communicator (shared-memory mailbox + CUDA IPC peer copies) instead of the torch.distributed callbacks
    comm = sb.ShmComm.from_process_group() if (len(sys.argv) > 3 and sys.argv[3] == "shm") else sb.ShardComm()
    for transcript in (sb.TRANSCRIPT_KECCAK, sb.TRANSCRIPT_BLAKE2B):
        plain = sb.create_proof(pk, instances, advice, seed, transcript)
        sharded = sb.create_proof(pk, instances, advice, seed, transcript, comm=comm)
        assert sharded == plain, f"rank {rank}: sharded proof differs from the single-GPU proof (transcript {transcript})"
    # the distributed four-step NTT (all-to-all between the passes, all-gather after): every rank ends with the local transform's result, bit for bit
    for ln in (16, 19, 20, 22):
        a = np.random.default_rng(ln).integers(0, 1 << 62, size=(1 << ln, 4), dtype=np.uint64)
        a[:, 3] &= np.uint64((1 << 61) - 1)   # < 2^253 < r: canonical residues
        w = fields.fr_to_mont(fields.omega(ln))
        want = sb.best_fft(a.copy(), w, ln, ctx)
        got = sb.best_fft_dist(a.copy(), w, ln, comm, ctx)
        assert (got == want).all(), f"rank {rank}: distributed NTT 2^{ln} differs from the local transform"
        winv, ninv = fields.fr_to_mont(pow(fields.omega(ln), -1, fields.FR_MODULUS)), fields.fr_to_mont(pow(1 << ln, -1, fields.FR_MODULUS))
        assert (sb.best_fft_dist(want.copy(), winv, ln, comm, ctx, scale=ninv) == a).all(), f"rank {rank}: distributed inverse NTT 2^{ln} (n^-1 folded in) is not the inverse"
    # timing (wall clock around the lock-step call, max over ranks)
    def timed(fn):
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        t = torch.tensor([(time.perf_counter() - t0) / reps * 1e3], device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_plain = timed(lambda: sb.create_proof(pk, instances, advice, seed, sb.TRANSCRIPT_KECCAK))
    t_shard = timed(lambda: sb.create_proof(pk, instances, advice, seed, sb.TRANSCRIPT_KECCAK, comm=comm))
    if rank == 0:
        print(f"SHARDED_OK k={k} world={world} plain_ms={t_plain:.2f} sharded_ms={t_shard:.2f}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
