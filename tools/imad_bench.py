"""Integer-multiply roofs of one B200 (developer tool): IMAD (low 32 bits), IMAD.HI, IMAD.WIDE -- warp instructions per cycle per SM sub-partition.

    python tools/imad_bench.py
"""
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import circuits_halo2_b200 as sb  # noqa: E402
from circuits_halo2_b200 import _lib  # noqa: E402

L = _lib.lib()
ctx = sb.Context(0)
props = torch.cuda.get_device_properties(0)
sm = props.multi_processor_count
clk = torch.cuda.clock_rate() * 1e6 if hasattr(torch.cuda, "clock_rate") else 1.965e9
blocks, threads, iters = sm * 8, 256, 4096
out = {}
for name in ("sb_bench_imad", "sb_bench_imad_hi", "sb_bench_imad_wide"):
    ms = ctypes.c_float()
    best = 1e9
    for _ in range(3):
        _lib.check(getattr(L, name)(ctx.handle, blocks, threads, iters, ctypes.byref(ms)), name)
        best = min(best, ms.value)
    ops = blocks * threads * iters * 8
    out[name] = {"ms": round(best, 4), "T_per_s": round(ops / (best * 1e-3) / 1e12, 3),
                 "cycles_per_warp_instr_per_smsp_at_1965MHz": round(1.965e9 * best * 1e-3 / (ops / 32 / (sm * 4)), 3)}
print(json.dumps(out))
