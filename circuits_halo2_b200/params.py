"""Mirror of `halo2_proofs::poly::kzg::commitment::ParamsKZG<Bn256>` (prover side).

Reference call sites: zk_prover/src/circuits/utils.rs:55 (`read`), :64 (`downsize`), :70 (`setup`),
and every `commit` / `commit_lagrange` inside keygen / create_proof (:75-76, :94-102).
File layout of `read`: u32 k || 2^k G1 (monomial) || 2^k G1 (Lagrange) ||
G2 || s.G2, raw Montgomery little-endian limbs.  The bases live on the GPU for the object's lifetime.
"""
from __future__ import annotations

import ctypes
import struct
from typing import Optional

import numpy as np

from . import _lib
from .context import Context, as_u64, default_context, ptr


class ParamsKZG:
    def __init__(self, k: int, g, g_lagrange, tail: bytes = b"", ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self._k = k
        self.n = 1 << k
        self.g = as_u64(g, 8)
        self.g_lagrange = as_u64(g_lagrange, 8)
        if self.g.shape[0] != self.n or self.g_lagrange.shape[0] != self.n:
            raise AssertionError("ParamsKZG: g / g_lagrange must hold 2^k points")
        self.tail = tail  # G2 || s.G2 (verifier side; untouched here)
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_srs_upload(self.ctx.handle, ctypes.c_uint32(k), ptr(self.g), ptr(self.g_lagrange), ctypes.byref(self._h)), "sb_srs_upload")

    @classmethod
    def from_device(cls, k: int, d_g: int, d_g_lagrange: int, ctx: Optional[Context] = None) -> "ParamsKZG":
        """Wrap bases already resident in HBM (device pointers, 2^k x 64 B each; caller keeps ownership)."""
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self._k, self.n = k, 1 << k
        self.g = self.g_lagrange = None
        self.tail = b""
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_srs_wrap_dev(self.ctx.handle, ctypes.c_uint32(k), ctypes.c_void_p(d_g), ctypes.c_void_p(d_g_lagrange), ctypes.byref(self._h)), "sb_srs_wrap_dev")
        return self

    @classmethod
    def setup(cls, k: int, tau: int, ctx: Optional[Context] = None, download: bool = True) -> "ParamsKZG":
        """`ParamsKZG::setup(k, rng)` with an explicit (UNSAFE, test-only) secret tau: g[i] = [tau^i] G and
        g_lagrange[i] = [L_i(tau)] G, computed entirely on the GPU.  `download=False` skips the host copies."""
        from . import fields
        self = cls.__new__(cls)
        self.ctx = ctx or default_context()
        self._k, self.n = k, 1 << k
        self.tail = b""
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_srs_setup_unsafe(self.ctx.handle, ctypes.c_uint32(k), ptr(fields.fr_to_mont(tau)), ctypes.byref(self._h)), "sb_srs_setup_unsafe")
        self.g = self.g_lagrange = None
        if download:
            self.g = np.zeros((self.n, 8), dtype=np.uint64)
            self.g_lagrange = np.zeros((self.n, 8), dtype=np.uint64)
            _lib.check(_lib.lib().sb_srs_download(self.ctx.handle, self._h, ptr(self.g), ptr(self.g_lagrange)), "sb_srs_download")
        return self

    @classmethod
    def read(cls, path: str, ctx: Optional[Context] = None) -> "ParamsKZG":
        with open(path, "rb") as f:
            data = f.read()
        (k,) = struct.unpack("<I", data[:4])
        n = 1 << k
        if len(data) < 4 + 128 * n:
            raise IOError("ParamsKZG::read: file too short")
        return cls(k, data[4 : 4 + 64 * n], data[4 + 64 * n : 4 + 128 * n], data[4 + 128 * n :], ctx)

    def k(self) -> int:
        return self._k

    def downsize(self, new_k: int, download: bool = True) -> "ParamsKZG":
        """`ParamsKZG::downsize` (utils.rs:62-66); returns the smaller params (halo2 mutates in place), Lagrange bases by a group-element iFFT on the GPU."""
        if new_k > self._k:
            raise AssertionError("downsize: k must not exceed the current k")
        out = ParamsKZG.__new__(ParamsKZG)
        out.ctx, out._k, out.n, out.tail = self.ctx, new_k, 1 << new_k, self.tail
        out._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_srs_downsize(self.ctx.handle, self._h, ctypes.c_uint32(new_k), ctypes.byref(out._h)), "sb_srs_downsize")
        out.g = out.g_lagrange = None
        if download:
            out.g = np.zeros((out.n, 8), dtype=np.uint64)
            out.g_lagrange = np.zeros((out.n, 8), dtype=np.uint64)
            _lib.check(_lib.lib().sb_srs_download(self.ctx.handle, out._h, ptr(out.g), ptr(out.g_lagrange)), "sb_srs_download")
        return out

    def precompute(self, bases: int = 3, window_bits: int = 0) -> "ParamsKZG":
        """Fixed-base window tables for `commit` (bit 0) / `commit_lagrange` (bit 1): 2^(c w) * P_i for every window w, so all
        windows share one bucket set and c can be 20-22 bits.  Same results, ceil(255 / c) x the base memory.  `ProvingKey`
        does this for its params on creation."""
        _lib.check(_lib.lib().sb_srs_precompute(self.ctx.handle, self._h, ctypes.c_int32(bases), ctypes.c_uint32(window_bits)), "sb_srs_precompute")
        return self

    @property
    def handle(self):
        return self._h

    def _commit(self, basis: int, scalars) -> np.ndarray:
        s = as_u64(scalars, 4)
        if s.shape[0] > self.n:
            raise AssertionError("commit: polynomial longer than the SRS")
        out = np.zeros(8, dtype=np.uint64)
        _lib.check(_lib.lib().sb_msm_g1(self.ctx.handle, self._h, ctypes.c_int32(basis), ptr(s), ctypes.c_size_t(s.shape[0]), ptr(out)), "sb_msm_g1")
        return out

    def commit(self, poly_coeffs) -> np.ndarray:
        """`ParamsKZG::commit`: MSM over the monomial bases; returns G1Affine (8,) uint64."""
        return self._commit(0, poly_coeffs)

    def commit_lagrange(self, poly_evals) -> np.ndarray:
        """`ParamsKZG::commit_lagrange`: MSM over the Lagrange bases."""
        return self._commit(1, poly_evals)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().sb_srs_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass
