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


def create_proof(pk: ProvingKey, instances: Sequence[int], advice, rng_seed: bytes, transcript: int = TRANSCRIPT_KECCAK) -> bytes:
    """One circuit instance.  instances: public inputs (python ints); advice: (A, n, 4) uint64 assigned advice columns;
    rng_seed: 32 bytes for ChaCha20Rng::from_seed.  Returns the proof bytes (transcript.finalize())."""
    if len(rng_seed) != 32:
        raise AssertionError("rng_seed must be 32 bytes")
    inst = np.concatenate([fields.fr_to_mont(int(v)) for v in instances]) if len(instances) else np.zeros(0, dtype=np.uint64)
    adv = np.ascontiguousarray(as_u64(advice, 4))
    n = 1 << pk.params.k()
    if adv.shape[0] != pk.cs["num_advice_columns"] * n:
        raise AssertionError("create_proof: advice must hold num_advice_columns x n cells")
    cap = 1 << 16
    out = np.zeros(cap, dtype=np.uint8)
    plen = ctypes.c_size_t()
    seed = np.frombuffer(rng_seed, dtype=np.uint8).copy()
    _lib.check(_lib.lib().sb_create_proof(pk.ctx.handle, pk.handle, ptr(inst), ctypes.c_size_t(len(instances)), ptr(adv), ptr(seed), ctypes.c_int32(transcript),
                                          ptr(out), ctypes.c_size_t(cap), ctypes.byref(plen)), "sb_create_proof")
    return out[: plen.value].tobytes()
