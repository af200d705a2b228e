"""ctypes binding of libsumma_b200.so (include/summa_b200.h).

There is NO CPU fallback: if the shared library is missing, or no CUDA device is visible when a
context is created, this module raises.  Nothing here imports `oracle/`.
"""
from __future__ import annotations

import ctypes
import os
import re
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "libsumma_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "summa_b200.h")

_lib = None


class SummaB200Error(RuntimeError):
    def __init__(self, status: int, what: str, detail: str):
        super().__init__(f"{what} failed with status {status}: {detail}")
        self.status = status


def build(force: bool = False) -> str:
    """Compile the CUDA extension for sm_100a (nvcc cross-compiles without a GPU)."""
    csrc = os.path.join(_HERE, "csrc")
    if force:
        subprocess.check_call(["make", "-s", "-C", csrc, "clean"])
    subprocess.check_call(["make", "-s", "-j4", "-C", csrc])
    return SO_PATH


def declared_symbols() -> list[str]:
    """Every function include/summa_b200.h declares (used by the CPU-side ABI test)."""
    with open(HEADER_PATH) as f:
        text = f.read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sb_[a-z0-9_]+)\s*\(", text)))


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(SO_PATH):
            raise ImportError(
                f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(libsumma_b200 has no CPU fallback)")
        L = ctypes.CDLL(SO_PATH)
        L.sb_last_error.restype = ctypes.c_char_p
        for name in declared_symbols():
            fn = getattr(L, name)
            if name != "sb_last_error":
                fn.restype = ctypes.c_int32
        _lib = L
    return _lib


def check(status: int, what: str) -> None:
    if status != 0:
        detail = lib().sb_last_error().decode(errors="replace")
        raise SummaB200Error(status, what, detail)
