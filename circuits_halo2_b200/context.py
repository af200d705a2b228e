"""One GPU context (`sb_ctx`): stream, scratch arena, NTT plan cache."""
from __future__ import annotations

import ctypes
from typing import Optional

import numpy as np

from . import _lib


def as_u64(a, cols: int) -> np.ndarray:
    """View `a` (bytes / ndarray) as a C-contiguous (n, cols) uint64 array without changing bytes."""
    if isinstance(a, (bytes, bytearray, memoryview)):
        arr = np.frombuffer(bytes(a), dtype=np.uint64)
    else:
        arr = np.ascontiguousarray(a)
        if arr.dtype != np.uint64:
            arr = arr.view(np.uint64)
    return arr.reshape(-1, cols)


def ptr(a: np.ndarray) -> ctypes.c_void_p:
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.c_void_p)


class Context:
    def __init__(self, device: int = 0):
        self._h = ctypes.c_void_p()
        L = _lib.lib()
        _lib.check(L.sb_ctx_create(ctypes.c_int32(device), ctypes.byref(self._h)), "sb_ctx_create")
        self.device = device

    @property
    def handle(self) -> ctypes.c_void_p:
        if not self._h:
            raise RuntimeError("context destroyed")
        return self._h

    def synchronize(self) -> None:
        _lib.check(_lib.lib().sb_ctx_synchronize(self.handle), "sb_ctx_synchronize")

    def set_blocking_sync(self, on: bool = True) -> None:
        """throughput mode: host waits sleep instead of spinning (use when many worker contexts share the host's cores)"""
        _lib.check(_lib.lib().sb_ctx_set_blocking_sync(self.handle, ctypes.c_int32(1 if on else 0)), "sb_ctx_set_blocking_sync")

    def stream(self) -> int:
        """the context's cudaStream_t (for CUDA-event timing by the caller)"""
        out = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_ctx_stream(self.handle, ctypes.byref(out)), "sb_ctx_stream")
        return out.value or 0

    def launch_count(self) -> int:
        out = ctypes.c_uint64()
        _lib.check(_lib.lib().sb_launch_count(self.handle, ctypes.byref(out)), "sb_launch_count")
        return out.value

    # raw device memory (for callers that do not bring torch)
    def alloc(self, nbytes: int) -> int:
        p = ctypes.c_void_p()
        _lib.check(_lib.lib().sb_dev_alloc(self.handle, ctypes.c_size_t(nbytes), ctypes.byref(p)), "sb_dev_alloc")
        return p.value

    def free(self, dptr: int) -> None:
        _lib.check(_lib.lib().sb_dev_free(self.handle, ctypes.c_void_p(dptr)), "sb_dev_free")

    def upload(self, dptr: int, host: np.ndarray) -> None:
        host = np.ascontiguousarray(host)
        _lib.check(_lib.lib().sb_dev_upload(self.handle, ctypes.c_void_p(dptr), ptr(host), ctypes.c_size_t(host.nbytes)), "sb_dev_upload")

    def download(self, dptr: int, nbytes: int) -> np.ndarray:
        out = np.empty(nbytes // 8, dtype=np.uint64)
        _lib.check(_lib.lib().sb_dev_download(self.handle, ptr(out), ctypes.c_void_p(dptr), ctypes.c_size_t(nbytes)), "sb_dev_download")
        return out

    def close(self) -> None:
        if self._h:
            _lib.lib().sb_ctx_destroy(self._h)
            self._h = ctypes.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_default: Optional[Context] = None


def default_context() -> Context:
    global _default
    if _default is None:
        import os
        _default = Context(int(os.environ.get("LOCAL_RANK", "0")) if os.environ.get("SB_USE_LOCAL_RANK") else 0)
    return _default
