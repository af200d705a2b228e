"""Mirror of `halo2_proofs::plonk::{ProvingKey, create_proof}` for KZG / SHPLONK over the C ABI.

Reference call sites: zk_prover/src/circuits/utils.rs:75-76 (keygen), :94-102 (`full_prover`,
Blake2b transcript), :171-178 (`gen_proof_solidity_calldata`, Keccak256 / EVM transcript).
The constraint system is the JSON description documented in include/summa_b200.h."""
from __future__ import annotations

import ctypes
import json
import struct
from typing import Optional, Sequence

import numpy as np

from . import _lib, fields
from .context import Context, as_u64, default_context, ptr
from .params import ParamsKZG

TRANSCRIPT_BLAKE2B = 0
TRANSCRIPT_KECCAK = 1


def seed_from_u64(state: int) -> bytes:
    """rand_core `SeedableRng::seed_from_u64`: the 32-byte ChaCha20Rng seed for a u64 (PCG32 expansion)."""
    mul, inc, mask = 6364136223846793005, 11634580027462260723, (1 << 64) - 1
    out = b""
    for _ in range(8):
        state = (state * mul + inc) & mask
        xorshifted = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
        rot = state >> 59
        out += struct.pack("<I", ((xorshifted >> rot) | (xorshifted << ((-rot) & 31))) & 0xFFFFFFFF)
    return out


class ProvingKey:
    """Device-resident `ProvingKey<G1Affine>`: fixed / permutation polynomials in all three forms + l_0, l_last, l_active."""

    def __init__(self, params: ParamsKZG, cs, fixed_values, sigma_values, transcript_repr: int, ctx: Optional[Context] = None):
        self.ctx = ctx or params.ctx
        self.params = params
        self.cs = json.loads(cs) if isinstance(cs, str) else cs
        text = json.dumps(self.cs).encode()
        n = 1 << params.k()
        fv = np.ascontiguousarray(as_u64(fixed_values, 4)).reshape(-1, 4)
        sv = np.ascontiguousarray(as_u64(sigma_values, 4)).reshape(-1, 4)
        self.num_fixed, self.num_sigma = self.cs["num_fixed_columns"], len(self.cs["permutation_columns"])
        if fv.shape[0] != self.num_fixed * n or sv.shape[0] != self.num_sigma * n:
            raise AssertionError("ProvingKey: fixed / sigma column sizes do not match the constraint system and k")
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_pk_create(self.ctx.handle, params.handle, ctypes.c_char_p(text), ctypes.c_uint32(params.k()), ptr(fv), ptr(sv),
                                           ptr(fields.fr_to_mont(transcript_repr)), ctypes.byref(self._h)), "sb_pk_create")

    @classmethod
    def from_sparse(cls, params: ParamsKZG, cs, fixed_cells, fixed_cell_values, perm_cells, transcript_repr: int, ctx: Optional[Context] = None) -> "ProvingKey":
        """Build the key from keygen's sparse output: fixed_cells (m, 2) uint32 (col, row) with fixed_cell_values (m, 4) uint64,
        perm_cells (p, 4) uint32 (col, row, to_col, to_row).  Everything dense is materialised on the GPU."""
        self = cls.__new__(cls)
        self.ctx = ctx or params.ctx
        self.params = params
        self.cs = json.loads(cs) if isinstance(cs, str) else cs
        text = json.dumps(self.cs).encode()
        self.num_fixed, self.num_sigma = self.cs["num_fixed_columns"], len(self.cs["permutation_columns"])
        fc = np.ascontiguousarray(fixed_cells, dtype=np.uint32).reshape(-1, 2)
        fvv = np.ascontiguousarray(as_u64(fixed_cell_values, 4)).reshape(-1, 4)
        pc = np.ascontiguousarray(perm_cells, dtype=np.uint32).reshape(-1, 4)
        if fc.shape[0] != fvv.shape[0]:
            raise AssertionError("ProvingKey.from_sparse: one value per fixed cell")
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_pk_create_sparse(self.ctx.handle, params.handle, ctypes.c_char_p(text), ctypes.c_uint32(params.k()), ptr(fc), ptr(fvv),
                                                  ctypes.c_size_t(fc.shape[0]), ptr(pc), ctypes.c_size_t(pc.shape[0]), ptr(fields.fr_to_mont(transcript_repr)),
                                                  ctypes.byref(self._h)), "sb_pk_create_sparse")
        return self

    @property
    def handle(self):
        return self._h

    def commitments(self):
        """(fixed_commitments, permutation_commitments) as (count, 8) uint64 G1Affine arrays (keygen_vk's output)."""
        f = np.zeros((self.num_fixed, 8), dtype=np.uint64)
        s = np.zeros((self.num_sigma, 8), dtype=np.uint64)
        _lib.check(_lib.lib().sb_pk_commitments(self._h, ptr(f), ptr(s)), "sb_pk_commitments")
        return f, s

    def __del__(self):
        try:
            if self._h:
                _lib.lib().sb_pk_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass


class _SbComm(ctypes.Structure):
    _AG_HOST = ctypes.CFUNCTYPE(ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
    _AG_DEV = ctypes.CFUNCTYPE(ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
    _A2A_DEV = ctypes.CFUNCTYPE(ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p)
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("user", ctypes.c_void_p), ("allgather_host", _AG_HOST), ("allgather_dev", _AG_DEV),
                ("alltoall_dev", _A2A_DEV)]


class _DevMem:
    """zero-copy view of library-owned device memory for torch (`__cuda_array_interface__`)"""

    def __init__(self, dptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (dptr, False), "version": 2}


def shard_range(n: int, rank: int, world: int):
    """base range of rank `rank` in an MSM of n points split over `world` GPUs (the last rank takes the remainder)"""
    per = n // world
    return rank * per, n if rank + 1 == world else (rank + 1) * per


class ShardComm:
    """`sb_comm` over a torch.distributed process group: one process per GPU, NCCL (device all-gather of the quotient's coset
    values over NVLink) and NCCL or gloo for the 64-byte partial commitments.  the quotient's j - 1 = 5 cosets are dealt round-robin to the ranks (`world` <= 8; ranks beyond the fifth idle in that stage)."""

    def __init__(self, group=None, device: Optional[int] = None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.group = torch, dist, group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        backend = str(dist.get_backend(group))
        self.host_on_cpu = "gloo" in backend
        self.device = device if device is not None else (torch.cuda.current_device() if torch.cuda.is_available() else None)
        self.error = None
        self._c = _SbComm(self.rank, self.world, None, _SbComm._AG_HOST(self._allgather_host), _SbComm._AG_DEV(self._allgather_dev),
                          _SbComm._A2A_DEV(self._alltoall_dev))

    def _allgather_host(self, _user, send, recv, nbytes):
        try:
            torch, dist = self.torch, self.dist
            mine = torch.frombuffer(bytearray(ctypes.string_at(send, nbytes)), dtype=torch.uint8)
            if self.host_on_cpu:
                out = torch.empty(self.world * nbytes, dtype=torch.uint8)
                dist.all_gather_into_tensor(out, mine, group=self.group)
            else:
                dev = torch.device("cuda", self.device)
                out_d = torch.empty(self.world * nbytes, dtype=torch.uint8, device=dev)
                dist.all_gather_into_tensor(out_d, mine.to(dev), group=self.group)
                out = out_d.cpu()
            ctypes.memmove(recv, out.numpy().ctypes.data, self.world * nbytes)
            return 0
        except Exception as e:  # never unwind into C
            self.error = e
            return 1

    def _allgather_dev(self, _user, d_buf, bytes_per_rank, stream):
        try:
            torch, dist = self.torch, self.dist
            full = torch.as_tensor(_DevMem(d_buf, bytes_per_rank * self.world), device=torch.device("cuda", self.device))
            ext = torch.cuda.ExternalStream(stream, device=torch.device("cuda", self.device)) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                dist.all_gather_into_tensor(full, full[self.rank * bytes_per_rank:(self.rank + 1) * bytes_per_rank], group=self.group)
            ext.synchronize()
            return 0
        except Exception as e:
            self.error = e
            return 1

    def _alltoall_dev(self, _user, d_send, d_recv, bytes_per_pair, stream):
        """NCCL all-to-all over NVLink (the transpose step of the distributed four-step NTT): block q of d_send -> rank q"""
        try:
            torch, dist = self.torch, self.dist
            dev = torch.device("cuda", self.device)
            send = torch.as_tensor(_DevMem(d_send, bytes_per_pair * self.world), device=dev)
            recv = torch.as_tensor(_DevMem(d_recv, bytes_per_pair * self.world), device=dev)
            ext = torch.cuda.ExternalStream(stream, device=dev) if stream else torch.cuda.current_stream()
            with torch.cuda.stream(ext):
                dist.all_to_all_single(recv, send, group=self.group)
            ext.synchronize()
            return 0
        except Exception as e:
            self.error = e
            return 1

    @property
    def struct(self):
        return self._c


class ShmComm:
    """`sb_comm` implemented INSIDE the library for the ranks of one box (sb_comm_shm_create): shared-memory mailbox for the small host records, CUDA IPC
    peer copies over NVLink for device buffers.  No torch / NCCL / Python on the data path.  `name`: same on every rank, unique per job."""

    def __init__(self, name: str, rank: int, world: int, ctx: Optional[Context] = None):
        self.rank, self.world, self.error = rank, world, None
        self._c = _SbComm()
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_comm_shm_create(ctx.handle if ctx is not None else None, ctypes.c_char_p(name.encode()), ctypes.c_int32(rank), ctypes.c_int32(world),
                                                 ctypes.byref(self._c), ctypes.byref(self._h)), "sb_comm_shm_create")

    @classmethod
    def from_process_group(cls, ctx: Optional[Context] = None, group=None) -> "ShmComm":
        """one communicator per torch.distributed process group: rank 0 picks a fresh name and broadcasts it"""
        import os
        import uuid
        import torch.distributed as dist
        box = [uuid.uuid4().hex[:16] if dist.get_rank(group) == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        return cls(f"{os.environ.get('MASTER_PORT', '0')}_{box[0]}", dist.get_rank(group), dist.get_world_size(group), ctx)

    def allgather_host(self, data: bytes) -> bytes:
        """the mailbox exchange on its own (tests)"""
        out = ctypes.create_string_buffer(len(data) * self.world)
        rc = self._c.allgather_host(self._c.user, data, out, len(data))
        if rc != 0:
            raise RuntimeError(_lib.lib().sb_last_error().decode(errors="replace"))
        return out.raw

    @property
    def struct(self):
        return self._c

    def close(self):
        if self._h:
            _lib.lib().sb_comm_shm_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LocalComm:
    """world = 1: the coset-sharded code path on a single GPU (all cosets owned, no exchange)"""

    def __init__(self):
        self.rank, self.world, self.error = 0, 1, None
        self._c = _SbComm(0, 1, None, _SbComm._AG_HOST(0), _SbComm._AG_DEV(0), _SbComm._A2A_DEV(0))

    @property
    def struct(self):
        return self._c


def _instances_mont(instances) -> np.ndarray:
    return np.concatenate([fields.fr_to_mont(int(v)) for v in instances]) if len(instances) else np.zeros(0, dtype=np.uint64)


def _finish(status: int, what: str, comm, out: np.ndarray, plen) -> bytes:
    if status != 0 and comm is not None and comm.error is not None:
        raise comm.error
    _lib.check(status, what)
    return out[: plen.value].tobytes()


def create_proof(pk: ProvingKey, instances: Sequence[int], advice, rng_seed: bytes, transcript: int = TRANSCRIPT_KECCAK, comm=None, ctx: Optional[Context] = None) -> bytes:
    """One circuit instance.  instances: public inputs (python ints); advice: (A, n, 4) uint64 assigned advice columns;
    rng_seed: 32 bytes for ChaCha20Rng::from_seed.  Returns the proof bytes (transcript.finalize()).
    comm (ShardComm): shard this proof over the ranks of a process group; every rank must call with the same arguments.
    ctx: run on another context of the key's device (its own stream and scratch; the key is read-only) -- see BatchProver."""
    if len(rng_seed) != 32:
        raise AssertionError("rng_seed must be 32 bytes")
    inst = _instances_mont(instances)
    adv = np.ascontiguousarray(as_u64(advice, 4))
    n = 1 << pk.params.k()
    if adv.shape[0] != pk.cs["num_advice_columns"] * n:
        raise AssertionError("create_proof: advice must hold num_advice_columns x n cells")
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.uint8)
    plen = ctypes.c_size_t()
    seed = np.frombuffer(rng_seed, dtype=np.uint8).copy()
    h = (ctx or pk.ctx).handle
    if comm is None:
        st = _lib.lib().sb_create_proof(h, pk.handle, ptr(inst), ctypes.c_size_t(len(instances)), ptr(adv), ptr(seed), ctypes.c_int32(transcript),
                                        ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
        return _finish(st, "sb_create_proof", None, out, plen)
    st = _lib.lib().sb_create_proof_sharded(h, pk.handle, ctypes.byref(comm.struct), ptr(inst), ctypes.c_size_t(len(instances)), ptr(adv), ptr(seed),
                                            ctypes.c_int32(transcript), ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
    return _finish(st, "sb_create_proof_sharded", comm, out, plen)


def create_proof_sparse(pk: ProvingKey, instances: Sequence[int], advice_cells, advice_values, rng_seed: bytes, transcript: int = TRANSCRIPT_KECCAK, comm=None,
                        ctx: Optional[Context] = None) -> bytes:
    """Same proof from the ASSIGNED advice cells only: advice_cells (m, 2) uint32 (column, row), advice_values (m, 4) uint64; every other
    cell is zero.  The bytes equal `create_proof` on the dense columns."""
    if len(rng_seed) != 32:
        raise AssertionError("rng_seed must be 32 bytes")
    inst = _instances_mont(instances)
    cells = np.ascontiguousarray(advice_cells, dtype=np.uint32).reshape(-1, 2)
    vals = np.ascontiguousarray(as_u64(advice_values, 4)).reshape(-1, 4)
    if cells.shape[0] != vals.shape[0]:
        raise AssertionError("create_proof_sparse: one value per advice cell")
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.uint8)
    plen = ctypes.c_size_t()
    seed = np.frombuffer(rng_seed, dtype=np.uint8).copy()
    h = (ctx or pk.ctx).handle
    if comm is None:
        st = _lib.lib().sb_create_proof_sparse(h, pk.handle, ptr(inst), ctypes.c_size_t(len(instances)), ptr(cells), ptr(vals), ctypes.c_size_t(cells.shape[0]), ptr(seed),
                                               ctypes.c_int32(transcript), ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
        return _finish(st, "sb_create_proof_sparse", None, out, plen)
    st = _lib.lib().sb_create_proof_sharded_sparse(h, pk.handle, ctypes.byref(comm.struct), ptr(inst), ctypes.c_size_t(len(instances)), ptr(cells), ptr(vals),
                                                   ctypes.c_size_t(cells.shape[0]), ptr(seed), ctypes.c_int32(transcript), ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
    return _finish(st, "sb_create_proof_sharded_sparse", comm, out, plen)


def create_proof_dev(pk: ProvingKey, instances: Sequence[int], d_advice: int, rng_seed: bytes, transcript: int = TRANSCRIPT_KECCAK, comm=None,
                     ctx: Optional[Context] = None) -> bytes:
    """Same proof with the advice columns already in device memory (`d_advice`: device pointer to A x n x 32 B; left untouched)."""
    if len(rng_seed) != 32:
        raise AssertionError("rng_seed must be 32 bytes")
    inst = _instances_mont(instances)
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.uint8)
    plen = ctypes.c_size_t()
    seed = np.frombuffer(rng_seed, dtype=np.uint8).copy()
    if comm is not None:
        st = _lib.lib().sb_create_proof_sharded_dev((ctx or pk.ctx).handle, pk.handle, ctypes.byref(comm.struct), ptr(inst), ctypes.c_size_t(len(instances)),
                                                    ctypes.c_void_p(d_advice), ptr(seed), ctypes.c_int32(transcript), ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
        return _finish(st, "sb_create_proof_sharded_dev", comm, out, plen)
    st = _lib.lib().sb_create_proof_dev((ctx or pk.ctx).handle, pk.handle, ptr(inst), ctypes.c_size_t(len(instances)), ctypes.c_void_p(d_advice), ptr(seed),
                                        ctypes.c_int32(transcript), ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen))
    return _finish(st, "sb_create_proof_dev", None, out, plen)


def mst_inclusion_witness(levels: int, n_currencies: int, k: int, preimages: np.ndarray, path_indices: np.ndarray, n_bytes: int = 8):
    """`MstInclusionCircuit::<LEVELS, N_CURRENCIES, N_BYTES>::init(merkle_proof)` + the witness side of `synthesize`
    (circuits/merkle_sum_tree.rs:86-103,228-520) for ONE Merkle proof in sb_mst_proofs' layout (MerkleSumTree.raw_proofs).
    Returns (instances as python ints, advice_cells (m, 2) uint32, advice_values (m, 4) uint64): the sparse witness create_proof_sparse takes."""
    pre = np.ascontiguousarray(preimages, dtype=np.uint64)
    path = np.ascontiguousarray(path_indices, dtype=np.uint8)
    n = ctypes.c_size_t()
    inst = np.zeros((2 + n_currencies, 4), dtype=np.uint64)
    L = _lib.lib()
    _lib.check(L.sb_mst_inclusion_witness(ctypes.c_uint32(levels), ctypes.c_uint32(n_currencies), ctypes.c_uint32(n_bytes), ctypes.c_uint32(k), ptr(pre), ptr(path),
                                          None, None, ctypes.c_size_t(0), ctypes.byref(n), ptr(inst)), "sb_mst_inclusion_witness")
    cells = np.zeros((n.value, 2), dtype=np.uint32)
    vals = np.zeros((n.value, 4), dtype=np.uint64)
    _lib.check(L.sb_mst_inclusion_witness(ctypes.c_uint32(levels), ctypes.c_uint32(n_currencies), ctypes.c_uint32(n_bytes), ctypes.c_uint32(k), ptr(pre), ptr(path),
                                          ptr(cells), ptr(vals), ctypes.c_size_t(n.value), ctypes.byref(n), ptr(inst)), "sb_mst_inclusion_witness")
    return [fields.fr_from_mont(inst[i]) for i in range(inst.shape[0])], cells, vals


class BatchProver:
    """Many independent proofs against ONE proving key (BASELINE configs[4]; the reference proves one user per
    `create_proof` call, backend/src/apis/round.rs:153-174, so a batch is many calls with a shared key).  Each worker thread
    owns a context (stream + scratch) on the key's GPU; the key and the SRS stay resident and are shared read-only, so the
    small kernels of different proofs overlap on the device and the host-side transcript work runs on several cores."""

    def __init__(self, pk: ProvingKey, workers: int = 4, blocking_sync: bool = False):
        from concurrent.futures import ThreadPoolExecutor
        import queue
        self.pk = pk
        self.contexts = [Context(pk.ctx.device) for _ in range(workers)]
        for c in self.contexts:
            c.set_blocking_sync(blocking_sync)   # workers sleep while their GPU work runs: the host cores go to transcripts and witness generation
        self._free = queue.SimpleQueue()
        for c in self.contexts:
            self._free.put(c)
        self._pool = ThreadPoolExecutor(max_workers=workers)

    def _one(self, job):
        instances, advice, seed, transcript = job
        c = self._free.get()
        try:
            return create_proof(self.pk, instances, advice, seed, transcript, ctx=c)
        finally:
            self._free.put(c)

    def prove_many(self, jobs):
        """jobs: iterable of (instances, advice, rng_seed, transcript); returns the proofs in job order."""
        return list(self._pool.map(self._one, jobs))

    def _one_sparse(self, job):
        instances, cells, values, seed, transcript = job
        c = self._free.get()
        try:
            return create_proof_sparse(self.pk, instances, cells, values, seed, transcript, ctx=c)
        finally:
            self._free.put(c)

    def _one_user(self, job):
        levels, n_cur, k, pre, path, seed, transcript = job
        c = self._free.get()
        try:
            inst, cells, vals = mst_inclusion_witness(levels, n_cur, k, pre, path)
            return create_proof_sparse(self.pk, inst, cells, vals, seed, transcript, ctx=c)
        finally:
            self._free.put(c)

    def prove_users(self, tree, indices, seeds, transcript: int = TRANSCRIPT_KECCAK):
        """BASELINE configs[4] / backend/src/apis/round.rs:153-174 for many users of one tree: Merkle proofs gathered on the GPU in one launch
        (`Tree::generate_proof`), then per user, on the worker threads: witness generation (`MstInclusionCircuit::init` + synthesize) and
        create_proof against the shared resident key.  Returns the proofs in `indices` order."""
        pre, path = tree.raw_proofs(indices)
        k = self.pk.params.k()
        jobs = [(tree.depth(), tree.n_currencies, k, pre[j], path[j], seeds[j], transcript) for j in range(len(indices))]
        return list(self._pool.map(self._one_user, jobs))

    def prove_many_sparse(self, jobs):
        """jobs: iterable of (instances, advice_cells, advice_values, rng_seed, transcript): the sparse-witness entry, one user per job."""
        return list(self._pool.map(self._one_sparse, jobs))

    def close(self):
        self._pool.shutdown(wait=True)
        for c in self.contexts:
            c.close()
        self.contexts = []
