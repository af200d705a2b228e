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
    def setup(cls, k: int, tau: int, ctx: Optional[Context] = None) -> "ParamsKZG":
        """`ParamsKZG::setup(k, rng)` with an explicit (UNSAFE, test-only) secret tau: g[i] = [tau^i] G and
        g_lagrange[i] = [L_i(tau)] G.  The 2 x 2^k fixed-base products run on the GPU; the scalars are host
        integers (setup is not on the proving path)."""
        from . import fields
        ctx = ctx or default_context()
        r, n = fields.FR_MODULUS, 1 << k
        w = fields.omega(k)
        mono, t = [], 1
        for _ in range(n):
            mono.append(t)
            t = t * tau % r
        tn = (pow(tau, n, r) - 1) % r
        ninv = pow(n, -1, r)
        # L_i(tau) = omega^i (tau^n - 1) / (n (tau - omega^i)), denominators inverted in one batch
        wi, dens, ws = 1, [], []
        for _ in range(n):
            ws.append(wi)
            dens.append((tau - wi) % r)
            wi = wi * w % r
        pref, acc = [], 1
        for dd in dens:
            pref.append(acc)
            acc = acc * dd % r
        inv = pow(acc, -1, r)
        lag = [0] * n
        for i in range(n - 1, -1, -1):
            lag[i] = ws[i] * tn % r * ninv % r * (inv * pref[i] % r) % r
            inv = inv * dens[i] % r
        scal = np.concatenate([fields.fr_to_mont(x) for x in mono + lag]).reshape(2 * n, 4)
        d_s = ctx.alloc(scal.nbytes)
        d_p = ctx.alloc(2 * n * 64)
        try:
            ctx.upload(d_s, scal)
            _lib.check(_lib.lib().sb_g1_fixed_base_mul_dev(ctx.handle, ctypes.c_void_p(d_s), ctypes.c_size_t(2 * n), ctypes.c_void_p(d_p), None), "sb_g1_fixed_base_mul_dev")
            ctx.synchronize()
            pts = ctx.download(d_p, 2 * n * 64).reshape(2 * n, 8)
        finally:
            ctx.free(d_s)
            ctx.free(d_p)
        return cls(k, pts[:n].copy(), pts[n:].copy(), b"", ctx)

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
