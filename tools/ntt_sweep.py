"""NTT variant sweep on one GPU (developer tool): sizes x {tile 2^11 / 2^12} x {fewest passes / 3 passes} x {full inter-pass tables / two-level}.
The knobs are the SB_NTT_* environment variables, read once per context (sb_ctx_create).  Prints one JSON line per variant.

    python tools/ntt_sweep.py [log_n ...]
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import circuits_halo2_b200 as sb  # noqa: E402
from circuits_halo2_b200 import _lib, fields  # noqa: E402
from circuits_halo2_b200.context import ptr  # noqa: E402

L = _lib.lib()
dev = torch.device("cuda", 0)
sizes = [int(x) for x in sys.argv[1:]] or [16, 18, 20, 22, 23, 24]
for ln in sizes:
    n = 1 << ln
    g = torch.Generator(device=dev)
    g.manual_seed(ln)
    a0 = torch.randint(-(1 << 63), (1 << 63) - 1, (n, 4), dtype=torch.int64, device=dev, generator=g)
    a0[:, 3] &= (1 << 61) - 1
    w = fields.fr_to_mont(fields.omega(ln))
    ref = None
    for tile in [int(t) for t in os.environ.get("SWEEP_TILES", "11,12").split(",")]:
        for passes in (0, 3):
            for tw_mb in (1024, 0):
                os.environ["SB_NTT_TILE"], os.environ["SB_NTT_PASSES"], os.environ["SB_NTT_TW_MB"] = str(tile), str(passes), str(tw_mb)
                ctx = sb.Context(0)
                st = ctypes.c_void_p(ctx.stream())
                ext = torch.cuda.ExternalStream(ctx.stream(), device=dev)
                a = a0.clone()
                torch.cuda.synchronize()
                run = lambda: _lib.check(L.sb_ntt_dev(ctx.handle, ctypes.c_void_p(a.data_ptr()), ptr(w), ctypes.c_uint32(ln), st), "sb_ntt_dev")
                l0 = ctx.launch_count()
                run()
                torch.cuda.synchronize()
                first = a.clone()
                npass = None
                if ref is None:
                    ref = first
                same = bool(torch.equal(first, ref))
                for _ in range(3):
                    run()
                l0 = ctx.launch_count()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                torch.cuda.synchronize()
                e0.record(ext)
                reps = 20
                for _ in range(reps):
                    run()
                e1.record(ext)
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / reps
                npass = (ctx.launch_count() - l0) // reps
                print(json.dumps({"log_n": ln, "tile": tile, "min_passes": passes, "tw_mb": tw_mb, "passes": npass, "ms": round(ms, 4), "gelem_per_s": round(n / ms / 1e6, 3),
                                  "same_result_as_first_variant": same}), flush=True)
                del a
                ctx.close()
