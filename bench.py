#!/usr/bin/env python
"""bench.py -- headline measurement of the hot path (BASELINE.json configs[1]).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--log-n 22] [--ntt-log-n 22]

A "step" is one BN254 G1 MSM (halo2 `best_multiexp` / `ParamsKZG::commit`) over 2^log_n synthetic
(random, valid) bases with uniform random scalars.  `value` times the device-resident call
(`sb_msm_g1_srs_dev`, inputs already in HBM); `e2e` times the reference-facing call
(`ParamsKZG.commit`, scalars in pinned HOST memory, H2D + D2H inside the timed region).
N > 1 (torchrun): the MSM is split by base range, one shard per rank ("weak": per-GPU work fixed),
the per-rank partial points are gathered and summed on the host (north_star), no other collective.

`--impl reference` times the restated halo2 CPU path (oracle/halo2_cpu.c: best_multiexp with
per-thread chunks, c = ceil(ln n)) with all host threads on a bounded sample of the same workload.
The Rust prover itself cannot be built in this image, so there is no oracle/_ref.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "msm_mpts_per_s"
UNIT = "Mpts/s"


def env_int(name, default):
    return int(os.environ.get(name, default))


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples SM clock and throttle reasons DURING the timed region through NVML (10 ms period)."""

    def __init__(self, index: int):
        self.index = index
        self.sm, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self.thread = None
        self.err = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                mask = int(get_reasons(self.h))
                for bit, nm in names.items():
                    if mask & bit:
                        self.reasons.add(nm)
            except Exception as e:  # pragma: no cover
                self.err = repr(e)
                break
            time.sleep(0.01)

    def stop(self):
        self._stop.set()
        if self.thread:
            self.thread.join(timeout=1)
        out = {"sm_mhz": statistics.median(self.sm) if self.sm else None, "sm_max_mhz": self.max_mhz,
               "reasons": sorted(self.reasons), "samples": len(self.sm)}
        if self.err:
            out["error"] = self.err
        return out


# ------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """Restated halo2 CPU `best_multiexp` on the host cores, bounded sample of the same workload."""
    if rank != 0:
        return
    import numpy as np
    from oracle import cpu
    cores = os.cpu_count() or 1
    log_s = min(args.log_n, args.cpu_sample_log_n)
    n = 1 << log_s
    bases = cpu.gen_bases(n, seed=1, threads=cores)
    scalars = cpu.random_fr(n, 2)
    for _ in range(args.warmup if args.warmup < 2 else 1):
        cpu.best_multiexp(scalars, bases, threads=cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu.best_multiexp(scalars, bases, threads=cores)
    dt = (time.perf_counter() - t0) / args.steps
    val = n / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 Montgomery (Fq/Fr, 254-bit)",
        "data": "synthetic", "config": {"workload": f"BN254 G1 MSM, 2^{args.log_n} points, uniform scalars (BASELINE configs[1])",
                                        "sample": f"2^{log_s} points per step"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"restated halo2 best_multiexp (oracle/halo2_cpu.c), 2^{log_s} points, {cores} threads"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--log-n", type=int, default=22, help="MSM size per GPU (2^log_n points)")
    ap.add_argument("--ntt-log-n", type=int, default=22, help="size of the NTT side measurement (0 = skip)")
    ap.add_argument("--no-tables", action="store_true", help="headline MSM without the fixed-base window tables (generic best_multiexp path)")
    ap.add_argument("--batch-k", type=int, default=13, help="k of the batched inclusion-proof side measurement (0 = skip)")
    ap.add_argument("--batch-proofs", type=int, default=64, help="proofs per GPU in the batch")
    ap.add_argument("--batch-workers", type=str, default="1,4,8", help="worker threads (contexts) per GPU to sweep")
    ap.add_argument("--mst-log-n", type=int, default=20, help="users (2^x) of the Merkle-sum-tree build side measurement (0 = skip)")
    ap.add_argument("--cpu-sample-log-n", type=int, default=22, help="size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--proof-k", type=str, default="17,20", help="comma-separated k values for the create_proof side measurement ('' = skip)")
    ap.add_argument("--dump-proof", type=str, default="", help="directory to write the last proof / vk commitments / instances to")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank, world, local = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist

    import circuits_halo2_b200 as sb
    from circuits_halo2_b200 import _lib
    from circuits_halo2_b200 import fields as fields_mod
    from circuits_halo2_b200.context import ptr

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = _lib.lib()
    ctx = sb.Context(local)
    dev = torch.device("cuda", local)
    # a dedicated (non-default) stream: kernels, copies and the timing events all live on it
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    st = ctypes.c_void_p(stream.cuda_stream)
    n = 1 << args.log_n

    def rand_fr_dev(count, seed):
        """uniform values < 2^253 (< r): any such value is a valid Montgomery residue"""
        g = torch.Generator(device=dev)
        g.manual_seed(seed)
        t = torch.randint(-(1 << 63), (1 << 63) - 1, (count, 4), dtype=torch.int64, device=dev, generator=g)
        t[:, 3] &= (1 << 61) - 1
        return t

    # ---- synthetic inputs, generated on the device by the product's own fixed-base kernel ----
    base_scalars = rand_fr_dev(n, 1000 + rank)
    bases = torch.empty((n, 8), dtype=torch.int64, device=dev)
    _lib.check(L.sb_g1_fixed_base_mul_dev(ctx.handle, ctypes.c_void_p(base_scalars.data_ptr()), ctypes.c_size_t(n),
                                          ctypes.c_void_p(bases.data_ptr()), st), "sb_g1_fixed_base_mul_dev")
    torch.cuda.synchronize()
    del base_scalars
    scalars = rand_fr_dev(n, 2000 + rank)
    host_scalars = torch.empty((n, 4), dtype=torch.int64).pin_memory()
    host_scalars.copy_(scalars)
    torch.cuda.synchronize()
    # SRS handle over the device-generated bases (one array serves as both bases of the handle)
    params = sb.ParamsKZG.from_device(args.log_n, bases.data_ptr(), bases.data_ptr(), ctx)
    plain_params = sb.ParamsKZG.from_device(args.log_n, bases.data_ptr(), bases.data_ptr(), ctx)  # same bases, no tables: the generic best_multiexp path
    if not args.no_tables:
        t0 = time.perf_counter()
        params.precompute(1)
        torch.cuda.synchronize()
        t_tables = time.perf_counter() - t0

    out = np.zeros(8, dtype=np.uint64)
    gather_buf = torch.zeros((world, 8), dtype=torch.int64, device=dev) if world > 1 else None

    def combine(local_out):
        """N > 1: gather the per-rank partial points and add them on the host (north_star)."""
        if world == 1:
            return local_out
        mine = torch.from_numpy(local_out.view(np.int64)).to(dev)
        dist.all_gather_into_tensor(gather_buf, mine)
        parts = gather_buf.cpu().numpy().view(np.uint64)
        res = np.zeros(8, dtype=np.uint64)
        _lib.check(L.sb_g1_sum_affine(ptr(np.ascontiguousarray(parts)), ctypes.c_size_t(world), ptr(res)), "sb_g1_sum_affine")
        return res

    def step_resident():
        _lib.check(L.sb_msm_g1_srs_dev(ctx.handle, params.handle, ctypes.c_int32(0), ctypes.c_void_p(scalars.data_ptr()), ctypes.c_size_t(n), ptr(out), st), "sb_msm_g1_srs_dev")
        return combine(out)

    hs_np = host_scalars.numpy().view(np.uint64)

    def step_e2e():
        return combine(params.commit(hs_np))  # H2D of the scalars + kernels + D2H of the window sums inside

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        for _ in range(steps):
            res = fn()
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, res

    for _ in range(max(args.warmup, 3)):
        r_res = step_resident()
    for _ in range(2):
        r_e2e = step_e2e()
    assert (r_res == r_e2e).all(), "resident and host-buffer paths disagree"

    # phase split of one MSM (CUDA events inside the library, on the launching stream)
    phase = (ctypes.c_float * 5)()
    shape = (ctypes.c_uint32 * 4)()

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count()
    phases = []
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(stream)
    for _ in range(args.steps):
        step_resident()
        L.sb_msm_phase_times(ctx.handle, phase, shape)
        phases.append(list(phase))
    e1.record(stream)
    barrier()
    ms_total = e0.elapsed_time(e1)
    launches = ctx.launch_count() - l0
    if world > 1:
        t = torch.tensor([ms_total], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6

    # the generic path (arbitrary bases, no precomputation: halo2 `best_multiexp(coeffs, bases)`), same inputs, same result
    def step_plain():
        _lib.check(L.sb_msm_g1_srs_dev(ctx.handle, plain_params.handle, ctypes.c_int32(0), ctypes.c_void_p(scalars.data_ptr()), ctypes.c_size_t(n), ptr(out), st), "sb_msm_g1_srs_dev")
        return combine(out)
    r_plain = step_plain()
    assert (r_plain == r_res).all(), "table and table-free MSM disagree"
    ms_plain, _ = timed(step_plain, args.steps)
    ms_plain /= args.steps
    L.sb_msm_phase_times(ctx.handle, phase, shape)
    plain_info = {"ms_per_step": ms_plain, "mpts_per_s": world * n / (ms_plain * 1e-3) / 1e6, "window_bits": int(shape[0]), "windows": int(shape[1]),
                  "phases_ms": {"recode_sort": phase[0], "reduce_level1": phase[1], "reduce_levels_ge2": phase[2], "bucket_reduce": phase[3], "device_total": phase[4]}}
    step_resident()  # leave the table path's shape in the context for the roofline below
    L.sb_msm_phase_times(ctx.handle, phase, shape)

    ms_e2e, _ = timed(step_e2e, args.steps)
    ms_e2e /= args.steps
    e2e_val = world * n / (ms_e2e * 1e-3) / 1e6

    # ---- roofline of the dominant kernel (msm_reduce_first_kernel: level-1 bucket accumulation) ----
    c, W, L1, seg = [int(x) for x in shape]
    k_ms = statistics.mean(p[1] for p in phases)
    mean_phase = [statistics.mean(p[i] for p in phases) for i in range(5)]
    # integer roof, measured live with the library's micro-kernels (MEASURED_PEAKS.json has no integer peak)
    sm = torch.cuda.get_device_properties(local).multi_processor_count
    ms_f = ctypes.c_float()
    blocks, threads, iters = sm * 8, 256, 4096
    _lib.check(L.sb_bench_imad(ctx.handle, blocks, threads, iters, ctypes.byref(ms_f)), "sb_bench_imad")
    imad_peak = blocks * threads * iters * 8 / (ms_f.value * 1e-3) / 1e12
    _lib.check(L.sb_bench_imad_wide(ctx.handle, blocks, threads, iters, ctypes.byref(ms_f)), "sb_bench_imad_wide")
    imadw_peak = blocks * threads * iters * 8 / (ms_f.value * 1e-3) / 1e12
    _lib.check(L.sb_bench_field_mul(ctx.handle, blocks, 128, 512, 1, ctypes.byref(ms_f)), "sb_bench_field_mul")
    fmul_peak = blocks * 128 * 512 * 4 / (ms_f.value * 1e-3) / 1e9  # G field-mul/s
    # algorithmic work: W windows x one mixed addition (10 field products) x 136 wide multiply-adds
    alg_imad = n * W * 10 * 136
    achieved = alg_imad / (k_ms * 1e-3) / 1e12
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    alg_bytes = n * W * (64 + 8) + n * 32  # gathered base + sorted (key,val) per digit, + scalar
    # DRAM traffic of the dominant kernel: one `ncu --set full` capture of this same workload, committed under profiles/
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r01_traffic.json")))
        if tj.get("log_n") == args.log_n and tj.get("fixed_base_tables") == (not args.no_tables):
            traffic, traffic_src = tj["msm_reduce_level1_dram_bytes_per_launch"], tj.get("source")
    except Exception:
        pass
    roofline = {
        "kernel": "msm_reduce_first_kernel (level-1 bucket accumulation)",
        "bound": "imad", "achieved": achieved, "peak": imadw_peak, "unit": "T wide-IMAD/s", "frac": achieved / imadw_peak,
        "peak_source": "measured live: sb_bench_imad_wide (8 independent IMAD.WIDE chains/thread)",
        "imad32_peak": imad_peak, "field_mul_peak_G_per_s": fmul_peak,
        "field_mul_frac": (n * W * 10 / (k_ms * 1e-3) / 1e9) / fmul_peak,
        "kernel_ms": k_ms, "kernel_share_of_step": k_ms / ms_step,
        "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"},
        "traffic": traffic, "traffic_source": traffic_src,
        "phases_ms": {"recode_sort": mean_phase[0], "reduce_level1": mean_phase[1], "reduce_levels_ge2": mean_phase[2],
                      "bucket_reduce": mean_phase[3], "device_total": mean_phase[4]},
    }

    # ---- NTT side measurement (same configs[1] sweep; not the headline value) ----
    # every rank transforms its own column (independent columns go to different GPUs: no exchange); aggregate = world x n / max time
    ntt = None
    if args.ntt_log_n:
        from circuits_halo2_b200 import fields
        ln = args.ntt_log_n
        nn = 1 << ln
        a = rand_fr_dev(nn, 3000 + rank)
        w = fields.fr_to_mont(fields.omega(ln))
        for _ in range(3):
            _lib.check(L.sb_ntt_dev(ctx.handle, ctypes.c_void_p(a.data_ptr()), ptr(w), ctypes.c_uint32(ln), st), "sb_ntt_dev")
        barrier()
        e0.record(stream)
        for _ in range(args.steps):
            _lib.check(L.sb_ntt_dev(ctx.handle, ctypes.c_void_p(a.data_ptr()), ptr(w), ctypes.c_uint32(ln), st), "sb_ntt_dev")
        e1.record(stream)
        barrier()
        t_ntt = e0.elapsed_time(e1) / args.steps
        if world > 1:
            t = torch.tensor([t_ntt], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            t_ntt = float(t.item())
        passes = 1 if ln <= 11 else -(-ln // 8)
        ntt = {"log_n": ln, "n_gpus": world, "ms": t_ntt, "melem_per_s": world * nn / (t_ntt * 1e-3) / 1e6, "passes": passes,
               "gb_per_s_actual_passes": world * 64 * nn * passes / (t_ntt * 1e-3) / 1e9,
               "gb_per_s_survey_def": world * 64 * nn * (-(-ln // 12)) / (t_ntt * 1e-3) / 1e9,
               "hbm_frac_actual": 64 * nn * passes / (t_ntt * 1e-3) / 1e9 / hbm_peak,
               "imad_frac": 68 * nn * ln / (t_ntt * 1e-3) / 1e12 / imadw_peak,
               "parallelism": "one independent column per GPU (replicas, no collective)"}
        del a

    # ---- Merkle-sum-tree build (BASELINE configs[2]'s 2^20-user snapshot): Keccak usernames + Poseidon tree on the device ----
    mst = None
    if rank == 0 and args.mst_log_n:
        nm = 1 << args.mst_log_n
        rng_m = np.random.default_rng(20)
        bal = rng_m.integers(0, 1 << 40, size=(nm, 2), dtype=np.uint64)
        names = [b"user_%d" % i for i in range(nm)]
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            tree = sb.MerkleSumTree.from_arrays(names, bal, ctx=ctx)
            wall = time.perf_counter() - t0
            best = tree.build_ms if best is None else min(best, tree.build_ms)
            root = tree.root()
            tree.close()
        assert root.balances == [int(bal[:, 0].astype(object).sum()), int(bal[:, 1].astype(object).sum())], "MST root balances != column sums"
        perms = nm * 3 + (nm - 1) * 4  # Poseidon permutations: leaf = N_CURRENCIES + 1, middle = N_CURRENCIES + 2 (N_CURRENCIES = 2)
        mst = {"users": nm, "currencies": 2, "device_ms": best, "wall_ms_incl_host_packing": wall * 1e3, "musers_per_s": nm / (best * 1e-3) / 1e6,
               "poseidon_permutations": perms, "G_field_mul_per_s": perms * 417 / (best * 1e-3) / 1e9, "field_mul_frac_of_peak": perms * 417 / (best * 1e-3) / 1e9 / fmul_peak}

    # ---- create_proof side measurement: the reference circuit MstInclusionCircuit<4,2,8> (entry_16.csv, user 0) at k = proof_k ----
    # N = 1: one GPU.  N > 1: the SAME proof sharded over the N ranks (sb_create_proof_sharded: MSMs by base range, evaluate_h / coset
    # NTTs by cosets of the extended domain); every rank builds the same key, the proof bytes must equal the single-GPU proof.
    proofs = []
    if args.proof_k:
        fx = np.load(os.path.join(ROOT, "tests", "golden", "mst_inclusion_assignment.npz"))
        cs_text = open(os.path.join(ROOT, "tests", "golden", "mst_inclusion_cs.json")).read()
        stage_names = ["advice_commit", "lookup_permute_commit", "permutation_product", "lookup_product", "random_poly", "coset_ntt", "evaluate_h",
                       "quotient_commit", "evaluations", "shplonk"]
        comm = sb.ShardComm(device=local) if world > 1 else None

        def stages():
            stg = (ctypes.c_float * 12)()
            L.sb_last_proof_stages(ctx.handle, stg)
            return {nm: round(float(stg[i]), 3) for i, nm in enumerate(stage_names)}

        for pk_k in [int(x) for x in args.proof_k.split(",") if x]:
            nrow = 1 << pk_k
            t0 = time.perf_counter()
            kzg = sb.ParamsKZG.setup(pk_k, 0x5A110000 + pk_k, ctx, download=False)
            t_srs = time.perf_counter() - t0
            t0 = time.perf_counter()
            pkey = sb.ProvingKey.from_sparse(kzg, cs_text, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234, ctx)
            t_pk = time.perf_counter() - t0
            adv_host = torch.zeros((3, nrow, 4), dtype=torch.int64).pin_memory()
            adv_np = adv_host.numpy().view(np.uint64)
            cells = fx["advice_cells"]
            adv_np[cells[:, 0], cells[:, 1]] = fx["advice_values"]
            insts = [fields_mod.fr_from_mont(v) for v in fx["instances"]]
            seed = sb.seed_from_u64(42)
            reps = max(2, min(args.steps, 5))

            def run(c):
                for _ in range(2):
                    pr = sb.create_proof(pkey, insts, adv_np, seed, sb.TRANSCRIPT_KECCAK, comm=c)
                barrier()
                l0p = ctx.launch_count()
                t0 = time.perf_counter()
                for _ in range(reps):
                    pr2 = sb.create_proof(pkey, insts, adv_np, seed, sb.TRANSCRIPT_KECCAK, comm=c)
                torch.cuda.synchronize()
                ms = (time.perf_counter() - t0) / reps * 1e3
                if world > 1:
                    t = torch.tensor([ms], device=dev)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ms = float(t.item())
                assert pr2 == pr, "create_proof is not deterministic in the seed"
                return pr, ms, (ctx.launch_count() - l0p) // reps, stages()

            proof, ms_proof, launches_p, stg_plain = run(None)
            hms = ctypes.c_float()
            hprog = (ctypes.c_uint32 * 4)()
            L.sb_last_h_profile(ctx.handle, ctypes.byref(hms), hprog)
            ext_pts = nrow * 8
            rec = {"k": pk_k, "ms_per_proof": ms_proof, "transcript": "keccak256/evm", "proof_bytes": len(proof),
                   "launches_per_proof": launches_p, "h2d_bytes_per_proof": 3 * nrow * 32,
                   "setup_srs_s": t_srs, "keygen_pk_s": t_pk, "stages_ms": stg_plain,
                   "evaluate_h": {"ms": hms.value, "instructions": int(hprog[0]), "field_mul": int(hprog[1]), "field_addsub": int(hprog[2]), "live_slots": int(hprog[3]),
                                  "G_field_mul_per_s": ext_pts * int(hprog[1]) / (hms.value * 1e-3) / 1e9 if hms.value else None,
                                  "field_mul_frac_of_peak": (ext_pts * int(hprog[1]) / (hms.value * 1e-3) / 1e9) / fmul_peak if hms.value else None}}
            if world > 1:
                sproof, ms_sh, launches_s, stg_sh = run(comm)
                assert sproof == proof, "sharded proof differs from the single-GPU proof"
                rec["sharded"] = {"n_gpus": world, "ms_per_proof": ms_sh, "speedup_vs_1gpu": ms_proof / ms_sh, "launches_per_proof_per_rank": launches_s,
                                  "stages_ms_rank0": stg_sh, "proof_equals_single_gpu": True,
                                  "timing": "wall clock around the lock-step call, barrier + synchronize on both sides, max over ranks"}
            proofs.append(rec)
            if args.dump_proof and rank == 0:
                os.makedirs(args.dump_proof, exist_ok=True)
                fcom, scom = pkey.commitments()
                np.savez(os.path.join(args.dump_proof, f"proof_k{pk_k}.npz"), proof=np.frombuffer(proof, dtype=np.uint8), fixed_comms=fcom, sigma_comms=scom,
                         instances=fx["instances"], k=np.array([pk_k]), transcript_repr=np.array([0x1234]))
            del pkey, kzg, adv_host

    # ---- batched inclusion proofs (BASELINE configs[4]): many independent create_proof calls against one resident key; every rank
    #      is a replica proving its own share ("replicas only": no collective), proofs/s is the sum over ranks ----
    batched = None
    if args.batch_k:
        bk, nb_proofs = args.batch_k, args.batch_proofs
        # the circuit of configs[4]: MstInclusionCircuit<20, 2, 8> (a tree of 2^20 users), minimum k = 13; tests/golden/make_assignment_l20.py
        l20 = os.path.join(ROOT, "tests", "golden", "mst_inclusion_assignment_l20.npz")
        use_l20 = bk >= 13 and os.path.exists(l20)
        fx = np.load(l20 if use_l20 else os.path.join(ROOT, "tests", "golden", "mst_inclusion_assignment.npz"))
        cs_text = open(os.path.join(ROOT, "tests", "golden", "mst_inclusion_cs.json")).read()
        kzg = sb.ParamsKZG.setup(bk, 0x5A110000 + bk, ctx, download=False)
        pkey = sb.ProvingKey.from_sparse(kzg, cs_text, fx["fixed_cells"], fx["fixed_values"], fx["perm_cells"], 0x1234, ctx)
        adv_np = torch.zeros((3, 1 << bk, 4), dtype=torch.int64).pin_memory().numpy().view(np.uint64)
        adv_np[fx["advice_cells"][:, 0], fx["advice_cells"][:, 1]] = fx["advice_values"]
        insts = [fields_mod.fr_from_mont(v) for v in fx["instances"]]
        jobs = [(insts, adv_np, sb.seed_from_u64(1000 * rank + j), sb.TRANSCRIPT_KECCAK) for j in range(nb_proofs)]
        first = sb.create_proof(pkey, *jobs[0])
        if args.dump_proof and rank == 0:
            os.makedirs(args.dump_proof, exist_ok=True)
            fcom, scom = pkey.commitments()
            np.savez(os.path.join(args.dump_proof, f"proof_batch_k{bk}.npz"), proof=np.frombuffer(first, dtype=np.uint8), fixed_comms=fcom, sigma_comms=scom,
                     instances=fx["instances"], k=np.array([bk]), transcript_repr=np.array([0x1234]))
        sweep = {}
        for workers in [int(x) for x in args.batch_workers.split(",") if x]:
            bp = sb.BatchProver(pkey, workers)
            bp.prove_many(jobs[: 2 * workers])  # warm-up: scratch arenas and NTT plans of every context
            barrier()
            t0 = time.perf_counter()
            out = bp.prove_many(jobs)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([dt], device=dev)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                dt = float(t.item())
            assert out[0] == first and len(set(out)) == len(out), "batched proofs must equal the sequential ones and differ per seed"
            sweep[str(workers)] = world * nb_proofs / dt
            bp.close()
        batched = {"k": bk, "proofs_per_rank": nb_proofs, "n_gpus": world, "proofs_per_s_by_workers_per_gpu": sweep, "best_proofs_per_s": max(sweep.values()),
                   "host_cores": os.cpu_count(), "circuit": ("MstInclusionCircuit<20,2,8> (LEVELS = 20: a 2^20-user tree), one witness, one ChaCha20 seed per proof" if use_l20
                               else "MstInclusionCircuit<4,2,8> witness of entry_16.csv user 0, one ChaCha20 seed per proof"),
                   "timing": "wall clock over the whole batch, barrier + synchronize on both sides, max over ranks"}
        del pkey, kzg

    # ---- CPU baseline (rank 0, bounded sample of the same workload) ----
    cpu_baseline = None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import cpu
        cores = os.cpu_count() or 1
        log_s = min(args.log_n, args.cpu_sample_log_n)
        ns = 1 << log_s
        hb = bases[:ns].cpu().numpy().view(np.uint64)
        hsc = np.ascontiguousarray(hs_np[:ns])
        cpu.best_multiexp(hsc[: ns // 4], hb[: ns // 4], threads=cores)  # warm-up
        t0 = time.perf_counter()
        ref = cpu.best_multiexp(hsc, hb, threads=cores)
        dt = time.perf_counter() - t0
        # the sample doubles as a parity check of the benchmarked inputs
        chk = np.zeros(8, dtype=np.uint64)
        d_sc = scalars[:ns].contiguous()
        _lib.check(L.sb_msm_g1_dev(ctx.handle, ctypes.c_void_p(bases.data_ptr()), ctypes.c_void_p(d_sc.data_ptr()), ctypes.c_size_t(ns), ptr(chk), st), "sb_msm_g1_dev")
        assert (chk == ref).all(), "GPU MSM != CPU oracle on the benchmark inputs"
        cpu_baseline = {"value": ns / dt / 1e6, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": f"restated halo2 best_multiexp (oracle/halo2_cpu.c), first 2^{log_s} of the benchmark's points, {cores} threads, {dt:.2f} s; result equals the GPU's"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x8 Montgomery (Fq/Fr, 254-bit)", "data": "synthetic",
            "config": {"workload": f"BN254 G1 MSM, 2^{args.log_n} points per GPU, uniform scalars (BASELINE configs[1])",
                       "window_bits": c, "windows": W, "level1_chunk": L1, "parallelism": f"base-range split x{world}, host fold",
                       "fixed_base_tables": (None if args.no_tables else {"entries_per_base": W, "bytes": W * n * 64, "build_s": t_tables,
                                                                          "note": "ParamsKZG bases are fixed: 2^(c w) P_i precomputed once, all windows share one bucket set"}),
                       "l2": "inputs (bases+scalars+sorted digits) exceed the 126 MB L2"},
            "e2e": {"value": e2e_val, "unit": UNIT, "ms_per_step": ms_e2e, "h2d_bytes_per_step": 32 * n, "d2h_bytes_per_step": 128 * W,
                    "api": "ParamsKZG.commit(host scalars) -> sb_msm_g1 (pinned host scalars, SRS resident like the reference's ParamsKZG)"},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": cpu_baseline,
            "extra": {"generic_best_multiexp_no_tables": plain_info, "ntt": ntt, "merkle_sum_tree": mst, "create_proof": proofs, "batched_inclusion_proofs": batched},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
