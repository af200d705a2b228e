"""BASELINE configs[1] sweep on one B200: BN254 G1 MSM (ParamsKZG.commit with fixed-base tables, and the generic best_multiexp path) for
2^16 .. 2^24 points and four scalar distributions (U uniform, Z 99 % zeros, C one value on 90 % of the rows, S 8-bit values),
and the Fr NTT (forward, natural order both sides) at the same sizes.  Device-resident inputs, CUDA events, 3 warm-ups, mean of 5.

    gpurun -- python tools/sweep.py > gpurun_out/sweep.json
"""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

import circuits_halo2_b200 as sb  # noqa: E402
from circuits_halo2_b200 import _lib, fields  # noqa: E402
from circuits_halo2_b200.context import ptr  # noqa: E402


def main():
    L = _lib.lib()
    ctx = sb.Context(0)
    dev = torch.device("cuda", 0)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = ctypes.c_void_p(stream.cuda_stream)
    g = torch.Generator(device=dev)

    def rand_fr(count, seed):
        g.manual_seed(seed)
        t = torch.randint(-(1 << 63), (1 << 63) - 1, (count, 4), dtype=torch.int64, device=dev, generator=g)
        t[:, 3] &= (1 << 61) - 1
        return t

    def scalars(kind, n, seed):
        s = rand_fr(n, seed)
        g.manual_seed(seed + 1)
        u = torch.rand(n, device=dev, generator=g)
        if kind == "Z":
            s[u < 0.99] = 0
        elif kind == "C":
            s[u < 0.9] = s[0].clone()
        elif kind == "S":
            vals = torch.stack([torch.from_numpy(fields.fr_to_mont(v).view(np.int64)) for v in range(256)]).to(dev)  # Montgomery forms of 0..255
            s = vals[(u * 256).long().clamp(0, 255)].contiguous()
        return s

    def timed(fn, reps=5):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(reps):
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    out = np.zeros(8, dtype=np.uint64)
    for log_n in (16, 18, 20, 22, 24):
        n = 1 << log_n
        bs = rand_fr(n, 1)
        bases = torch.empty((n, 8), dtype=torch.int64, device=dev)
        _lib.check(L.sb_g1_fixed_base_mul_dev(ctx.handle, ctypes.c_void_p(bs.data_ptr()), ctypes.c_size_t(n), ctypes.c_void_p(bases.data_ptr()), st), "bases")
        torch.cuda.synchronize()
        del bs
        plain = sb.ParamsKZG.from_device(log_n, bases.data_ptr(), bases.data_ptr(), ctx)
        tab = sb.ParamsKZG.from_device(log_n, bases.data_ptr(), bases.data_ptr(), ctx).precompute(1)
        rec = {"log_n": log_n, "msm": {}}
        for kind in "UZCS":
            sc = scalars(kind, n, 100 + log_n)
            res = {}
            for name, params in (("tables", tab), ("generic", plain)):
                def run():
                    _lib.check(L.sb_msm_g1_srs_dev(ctx.handle, params.handle, ctypes.c_int32(0), ctypes.c_void_p(sc.data_ptr()), ctypes.c_size_t(n), ptr(out), st), "msm")
                ms = timed(run)
                res[name] = {"ms": round(ms, 4), "mpts_per_s": round(n / ms / 1e3, 1), "result": out.copy().tolist()}
            assert res["tables"]["result"] == res["generic"]["result"], (log_n, kind)
            for v in res.values():
                del v["result"]
            rec["msm"][kind] = res
            del sc
        a = rand_fr(n, 7)
        w = fields.fr_to_mont(fields.omega(log_n))
        ms = timed(lambda: _lib.check(L.sb_ntt_dev(ctx.handle, ctypes.c_void_p(a.data_ptr()), ptr(w), ctypes.c_uint32(log_n), st), "ntt"))
        passes = 1 if log_n <= 11 else -(-log_n // 8)
        rec["ntt"] = {"ms": round(ms, 4), "gelem_per_s": round(n / ms / 1e6, 3), "gb_per_s": round(64 * n * passes / ms / 1e6, 1), "passes": passes}
        print(json.dumps(rec), flush=True)
        del a, bases, plain, tab
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
