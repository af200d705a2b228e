"""Mirror of `halo2_proofs::arithmetic::{best_multiexp, best_fft}`.

Reference call sites: zk_prover/src/circuits/utils.rs:75-76,94-102,171-178 (through
ParamsKZG::commit* and EvaluationDomain).  Same argument meaning and error behaviour:
length mismatches raise (the Rust asserts), results are exact group / field elements.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .context import Context, as_u64, default_context, ptr


def best_multiexp(coeffs, bases, ctx: Optional[Context] = None) -> np.ndarray:
    """sum_i coeffs[i] * bases[i].  Returns `G1` (Jacobian x, y, z with z = 1; identity (0, 1, 0))
    as a (12,) uint64 array in halo2curves layout."""
    ctx = ctx or default_context()
    c = as_u64(coeffs, 4)
    b = as_u64(bases, 8)
    if c.shape[0] != b.shape[0]:
        raise AssertionError("best_multiexp: coeffs.len() != bases.len()")
    out = np.zeros(12, dtype=np.uint64)
    _lib.check(_lib.lib().sb_best_multiexp(ctx.handle, ptr(c), ptr(b), ctypes.c_size_t(c.shape[0]), ptr(out)), "sb_best_multiexp")
    return out


def best_fft(a, omega, log_n: int, ctx: Optional[Context] = None) -> np.ndarray:
    """In-place radix-2 NTT with natural-order input and output; returns the (n, 4) array."""
    ctx = ctx or default_context()
    arr = as_u64(a, 4)
    if arr.shape[0] != (1 << log_n):
        raise AssertionError("best_fft: a.len() != 1 << log_n")
    if not arr.flags["WRITEABLE"]:
        arr = arr.copy()
    w = as_u64(omega, 4)
    _lib.check(_lib.lib().sb_best_fft(ctx.handle, ptr(arr), ptr(w), ctypes.c_uint32(log_n)), "sb_best_fft")
    return arr


def best_fft_dist(a, omega, log_n: int, comm, ctx: Optional[Context] = None, scale=None) -> np.ndarray:
    """`best_fft` as a distributed four-step NTT over the ranks of `comm` (ShardComm / LocalComm): every rank passes the SAME vector, does
    1 / world of both passes, and gets the whole transform back (all-to-all between the passes, all-gather at the end, over NVLink)."""
    ctx = ctx or default_context()
    arr = as_u64(a, 4)
    if arr.shape[0] != (1 << log_n):
        raise AssertionError("best_fft_dist: a.len() != 1 << log_n")
    w = as_u64(omega, 4)
    d = ctx.alloc(arr.nbytes)
    try:
        ctx.upload(d, arr)
        sc = ptr(as_u64(scale, 4)) if scale is not None else None
        st = _lib.lib().sb_ntt_dist(ctx.handle, ctypes.byref(comm.struct), ctypes.c_void_p(d), ptr(w), ctypes.c_uint32(log_n), sc, None)
        if st != 0 and getattr(comm, "error", None) is not None:
            raise comm.error
        _lib.check(st, "sb_ntt_dist")
        ctx.synchronize()
        return ctx.download(d, arr.nbytes).reshape(-1, 4)
    finally:
        ctx.free(d)
