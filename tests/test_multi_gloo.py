"""world_size-2 gloo tests (CPU) of the host side of the sharded paths: the ShardComm callbacks the C library
calls during sb_create_proof_sharded, the base-range partition, and the host fold of the per-rank partial commitments.
The per-rank partial MSMs are computed by the CPU oracle here (no GPU in this suite); the fold is the product's own
sb_g1_sum_affine."""
import ctypes
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    try:
        sys.path.insert(0, ROOT)
        import torch.distributed as dist
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        dist.init_process_group("gloo", rank=rank, world_size=world)
        import circuits_halo2_b200 as sb
        from circuits_halo2_b200 import _lib
        from circuits_halo2_b200.context import ptr
        from oracle import cpu
        comm = sb.ShardComm()
        assert (comm.rank, comm.world) == (rank, world) and comm.host_on_cpu
        # 1. the callback the library calls: gather 64-byte records in rank order
        send = (ctypes.c_uint8 * 64)(*([rank + 1] * 64))
        recv = (ctypes.c_uint8 * (64 * world))()
        rc = comm.struct.allgather_host(None, ctypes.cast(send, ctypes.c_void_p), ctypes.cast(recv, ctypes.c_void_p), 64)
        assert rc == 0 and bytes(recv) == b"".join(bytes([r + 1]) * 64 for r in range(world))
        # 2. base-range split of one MSM + host fold == the unsplit MSM
        n = 3001  # ragged: the last rank takes the remainder
        bases = cpu.gen_bases(n, seed=5, threads=2)
        scalars = cpu.random_fr(n, 6)
        lo, hi = sb.shard_range(n, rank, world)
        part = cpu.best_multiexp(scalars[lo:hi], bases[lo:hi], threads=2)
        assert part.shape == (8,)
        all_parts = (ctypes.c_uint8 * (64 * world))()
        pbytes = np.ascontiguousarray(part).tobytes()
        assert comm.struct.allgather_host(None, ctypes.cast(ctypes.c_char_p(pbytes), ctypes.c_void_p), ctypes.cast(all_parts, ctypes.c_void_p), 64) == 0
        total = np.zeros(8, dtype=np.uint64)
        parts = np.frombuffer(bytes(all_parts), dtype=np.uint64).copy()
        _lib.check(_lib.lib().sb_g1_sum_affine(ptr(parts), ctypes.c_size_t(world), ptr(total)), "sb_g1_sum_affine")
        full = cpu.best_multiexp(scalars, bases, threads=2)
        assert (total == full).all()
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))


def test_shard_ranges_cover_the_bases_exactly_once():
    sys.path.insert(0, ROOT)
    import circuits_halo2_b200 as sb
    for n in (1, 7, 8, 1000, 1 << 20, (1 << 20) + 5):
        for world in (1, 2, 4, 8):
            edges = [sb.shard_range(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))


def test_sharded_msm_host_fold_world2_gloo():
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctxm.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _shm_worker(rank, world, name, q):
    try:
        sys.path.insert(0, ROOT)
        import circuits_halo2_b200 as sb
        comm = sb.ShmComm(name, rank, world)          # host-only: no context, no GPU
        for it in range(3000):
            size = (16, 128, 1024, 8 * 128)[it % 4]
            mine = bytes([(rank * 37 + it + j) & 0xFF for j in range(size)])
            got = comm.allgather_host(mine)
            want = b"".join(bytes([(r * 37 + it + j) & 0xFF for j in range(size)]) for r in range(world))
            assert got == want, (rank, it)
        # a record larger than a mailbox slot is an error on every rank, not a hang
        try:
            comm.allgather_host(b"x" * (17 << 10))
            raise AssertionError("oversized record accepted")
        except RuntimeError:
            pass
        comm.close()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, traceback.format_exc() + repr(e)))


@pytest.mark.parametrize("world", [2, 4])
def test_shared_memory_mailbox_allgather(world):
    """sb_comm_shm (csrc/comm_shm.cu), host half: the ranks' small records (partial commitments, IPC handles) meet in a POSIX shared-memory mailbox.
    3000 back-to-back exchanges of changing sizes between `world` processes: every rank always reads every rank's record of the SAME generation."""
    import uuid
    import torch.multiprocessing as mp
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    name = "t" + uuid.uuid4().hex[:12]
    procs = [ctxm.Process(target=_shm_worker, args=(r, world, name, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(r, "ok") for r in range(world)], res
