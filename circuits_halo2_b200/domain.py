"""Mirror of `halo2_proofs::poly::EvaluationDomain<Fr>` over the C ABI."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib
from .context import Context, as_u64, default_context, ptr


class EvaluationDomain:
    """`EvaluationDomain::new(j, k)`: j = constraint-system degree, n = 2^k rows."""

    def __init__(self, j: int, k: int, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.j, self._k = j, k
        self._h = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_domain_create(self.ctx.handle, ctypes.c_uint32(j), ctypes.c_uint32(k), ctypes.byref(self._h)), "sb_domain_create")
        e = ctypes.c_uint32()
        _lib.check(_lib.lib().sb_domain_extended_k(self._h, ctypes.byref(e)), "sb_domain_extended_k")
        self._ext_k = e.value

    def k(self) -> int:
        return self._k

    def extended_k(self) -> int:
        return self._ext_k

    def extended_len(self) -> int:
        return 1 << self._ext_k

    def get_quotient_poly_degree(self) -> int:
        return self.j - 1

    @property
    def handle(self):
        return self._h

    def _inplace(self, fn, a, n):
        arr = as_u64(a, 4).copy()
        if arr.shape[0] != n:
            raise AssertionError("polynomial length does not match the domain")
        _lib.check(fn(self.ctx.handle, self._h, ptr(arr)), fn.__name__)
        return arr

    def lagrange_to_coeff(self, a) -> np.ndarray:
        return self._inplace(_lib.lib().sb_lagrange_to_coeff, a, 1 << self._k)

    def coeff_to_lagrange(self, a) -> np.ndarray:
        return self._inplace(_lib.lib().sb_coeff_to_lagrange, a, 1 << self._k)

    def coeff_to_extended(self, a) -> np.ndarray:
        c = as_u64(a, 4)
        if c.shape[0] != (1 << self._k):
            raise AssertionError("coeff_to_extended: a.len() != n")
        out = np.empty((1 << self._ext_k, 4), dtype=np.uint64)
        _lib.check(_lib.lib().sb_coeff_to_extended(self.ctx.handle, self._h, ptr(c), ptr(out)), "sb_coeff_to_extended")
        return out

    def extended_to_coeff(self, a) -> np.ndarray:
        e = as_u64(a, 4)
        if e.shape[0] != (1 << self._ext_k):
            raise AssertionError("extended_to_coeff: a.len() != extended_len")
        out = np.empty(((self.j - 1) << self._k, 4), dtype=np.uint64)
        _lib.check(_lib.lib().sb_extended_to_coeff(self.ctx.handle, self._h, ptr(e), ptr(out)), "sb_extended_to_coeff")
        return out

    def divide_by_vanishing_poly(self, a) -> np.ndarray:
        return self._inplace(_lib.lib().sb_divide_by_vanishing_poly, a, 1 << self._ext_k)

    def __del__(self):
        try:
            if self._h:
                _lib.lib().sb_domain_destroy(self._h)
                self._h = ctypes.c_void_p()
        except Exception:
            pass
