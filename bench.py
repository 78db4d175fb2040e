#!/usr/bin/env python
"""bench.py -- Gpoints/s of InterpGrid valgrad on B200 (BASELINE.json metric), one JSON line.

Headline workload (config.workload = "C2"): BASELINE configs[1] -- 3-D InterpCubic BCnan on a 256^3
Float64 grid (stored 258^3 after the cubic padding) incl. the interp_invert! prefilter at
construction, 1e7 uniform-random queries per GPU, valgrad (value + 3 gradient components).

A "step" is one pass of valgrad over one batch of 1e7 queries.
  value  device-resident: queries / outputs already in HBM; CUDA-event timed on the launch stream.
  e2e    the same step through the C-ABI with HOST buffers (pinned): H2D of the queries and D2H of
         value+gradient inside the timed region (gb_eval, mem=GB_HOST).  e2e_pageable repeats it with
         plain malloc'd numpy buffers (what a Julia Array is).
  roofline  the dominant kernel of the step is eval_brick_kernel<double,3,Cubic,nil,valgrad,...>; its
         average duration is measured live with CUDA events on the launch stream (gb_profile_*),
         algorithmic bytes = 56 B/query (24 in + 32 out) x 1e7 queries per launch ->
         achieved GB/s vs the measured HBM peak.  roofline.step_* gives the same figure for the whole
         step (binning sort + brick evaluation + un-permute), which is what `value` measures.
  contraction  the headline runs the library default, EXACT (the reference's arithmetic order,
         bit-identical to the oracle); extra.other_contraction repeats the step with GB_FAST
         (separable FMA, within north_star's 4 ulp of the magnitude scale -- tests/test_gpu_parity.py).
  cpu_baseline  the C++ restatement of the reference algorithm (oracle, "port") on the host cores.
  extra.configs  (N = 1) the other BASELINE configs on the same box, same timing method: C1 (1e6 and 1e8
         queries), C2 with tile-sorted queries, C3 (the whole 1e9-query batch in 1e8 slices, and the 8192^2
         prefilter), C4 (three boundary conditions), C5 (restrict and prolong chains), each with its
         roofline fraction, ncu DRAM traffic (profiles/traffic.json) and a bitwise parity check of a
         strided subsample against the oracle.
  extra.c3_sharded / extra.c5_slab  (N > 1) C3's 1e9 queries split over the ranks, and the C5 chain
         slab-sharded along dim 3 with the NCCL halo exchange (per-level ms, halo bytes, exchange wait).
Multi-GPU (torchrun, one process per GPU): the grid is prefiltered on rank 0 and the coefficient
array broadcast over NCCL; query batches shard with no data-path collective ("weak" scaling: 1e7
queries per GPU).
`--impl reference` times the reference's CPU algorithm (oracle port; Julia cannot run here).
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_GRID = 256
NQ = 10_000_000
BYTES_PER_QUERY = 56  # 3*8 coordinates in + (1+3)*8 out
SEED_GRID, SEED_Q = 2, 12
METRIC = "valgrad throughput, 3-D InterpCubic BCnan 256^3 Float64"


def host_cores():
    """Host threads this process may use -- NOT omp_get_max_threads(): torchrun exports OMP_NUM_THREADS=1."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def c2_config(world, contraction):
    """One dict for both arms (the driver compares them key by key)."""
    return {"workload": "C2", "grid": "256^3 f64 (stored 258^3)", "order": "InterpCubic", "bc": "BCnan", "outputs": "valgrad",
            "queries_per_gpu_per_step": NQ, "query_distribution": "uniform [1,256]^3, splitmix64", "contraction": contraction,
            "l2": "inputs larger than L2 (240 MB in + 320 MB out + 137 MB grid per step vs 126 MB L2)",
            "parallelism": f"query-sharded x{world}, coefficients replicated by NCCL broadcast"}


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def load_traffic():
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


class ClockSampler:
    """SM clock / throttle reasons during the timed region (pynvml; nvidia-smi values)."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz, self._stop = [], set(), None, threading.Event()
        self._t = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        while not self._stop.is_set():
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.001)  # the timed region is ~30 ms: sample densely

    def __enter__(self):
        if self.nv:
            self._t = threading.Thread(target=self._run, daemon=True)
            self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._t:
            self._t.join()

    def summary(self, window=None):
        """Median SM clock over the samples inside `window` = (t0, t1) (the timed region); the sampler runs
        from the warm-up to the end of the GPU work, so `samples_under_load` covers the whole busy phase."""
        allv = sorted(v for _, v in self.samples)
        inw = sorted(v for t, v in self.samples if window and window[0] <= t <= window[1])
        use = inw if inw else allv
        return {"sm_mhz": (use[len(use) // 2] if use else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(inw), "samples_under_load": len(allv), "sm_mhz_min_under_load": (allv[0] if allv else None)}


def run_reference(args):
    """The reference's own CPU algorithm (oracle port) on the host cores, same config/metric."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    import numpy as np

    from oracle import oracle as O

    n = N_GRID
    threads = host_cores()
    A = O.synth(np.float64, n ** 3, SEED_GRID).reshape((n, n, n), order="F")
    t0 = time.perf_counter()
    co = O.make_coefs(A, "nan", "cubic")
    t_pref = time.perf_counter() - t0
    sample = int(os.environ.get("GB_REF_SAMPLE", str(NQ)))  # the whole 1e7-query batch: ~2 s per step on 16 cores
    X = 1.0 + O.synth(np.float64, 3 * sample, SEED_Q).reshape(sample, 3) * (n - 1.0)
    for _ in range(max(1, min(args.warmup, 2))):
        O.evaluate(co, "nan", "cubic", X[: sample // 10], 3, nthreads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        O.evaluate(co, "nan", "cubic", X, 3, nthreads=threads)
    dt = (time.perf_counter() - t0) / args.steps
    val = sample / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "Gpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c2_config(args.gpus, "EXACT"),
        "note": "Julia cannot run in this image: this is the C++ restatement of the reference algorithm (oracle), OpenMP over queries on one host",
        "cpu_baseline": {"value": val, "unit": "Gpoints/s", "cores": threads, "kind": "port", "sample": f"{sample} of 1e7 queries per step, OpenMP over queries",
                         "prefilter_s": t_pref},
        "e2e": {"value": val, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


# ---- helpers for the per-config section -------------------------------------------------------------------------------
def _timed(torch, fn, reps):
    """best / median of `reps` CUDA-event timings after two warm-up calls (the batches are larger than L2)."""
    fn()
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[0], ts[len(ts) // 2]


def _row(workload, units, unit_name, bytes_per_unit, best_ms, med_ms, peak, parity, traffic=None, note=None, **more):
    gups = units / (med_ms * 1e-3) / 1e9
    r = {"workload": workload, "ms": med_ms, "ms_best": best_ms, "value": gups, "unit": "G%s/s" % unit_name,
         "roofline": {"bound": "hbm", "achieved": gups * bytes_per_unit, "peak": peak, "unit": "GB/s", "frac": gups * bytes_per_unit / peak,
                      "algorithmic_bytes_per_unit": bytes_per_unit, "traffic": traffic},
         "parity_vs_oracle": parity}
    if note:
        r["note"] = note
    r.update(more)
    return r


def extra_configs(gb, torch, np, O, peak, traffic, reps=5):
    """C1, C2 (sorted queries), C3, C4, C5 on one GPU; see the module docstring."""
    rows = []
    dev = torch.device("cuda", torch.cuda.current_device())

    def sub(t, idx):
        return t[idx].cpu().numpy()

    # ---- C1 ----------------------------------------------------------------------------------------------------
    n = 1024
    A = gb.jl_empty((n, n), torch.float64, dev)
    gb.synth_uniform(A, 1)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    co = G.coefs
    for nq in (1_000_000, 100_000_000):
        X = torch.empty((nq, 2), dtype=torch.float64, device=dev)
        gb.synth_uniform(X, 11, lo=1.0, hi=float(n))
        v = torch.empty(nq, dtype=torch.float64, device=dev)
        best, med = _timed(torch, lambda: gb.getindex_(v, G, X), reps)
        idx = torch.arange(0, nq, max(1, nq // 50_000), device=dev)
        ov, _, _, _ = O.evaluate(co, "reflect", "quadratic", sub(X, idx), 1)
        rows.append(_row("C1 2-D InterpQuadratic BCreflect 1024^2 f64 getindex, %.0e queries" % nq, nq, "points", 24, best, med, peak,
                         bool(np.array_equal(ov, sub(v, idx), equal_nan=True)), traffic.get("c1_plain_bytes_per_1e8") if nq == 100_000_000 else None))
        del X, v
    del G, A

    # ---- C2 with block-sorted queries (coherent callers: evaluated in place, no permutation) --------------------------------------
    n, nq = N_GRID, NQ
    A = gb.jl_empty((n, n, n), torch.float64, dev)
    gb.synth_uniform(A, SEED_GRID)
    G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
    X = torch.empty((nq, 3), dtype=torch.float64, device=dev)
    gb.synth_uniform(X, SEED_Q, lo=1.0, hi=float(n))
    fl = X.floor().to(torch.int64)
    key = ((fl[:, 2] >> 3) * 4096 + (fl[:, 1] >> 4) * 64 + (fl[:, 0] >> 4))  # 16 x 16 x 8-cell tiles, dim 1 fastest
    Xs = X[torch.argsort(key)].contiguous()
    del fl, key, X
    v = torch.empty(nq, dtype=torch.float64, device=dev)
    g = torch.empty((nq, 3), dtype=torch.float64, device=dev)
    co = G.coefs
    idx = torch.arange(0, nq, 200, device=dev)
    for fast in (False, True):
        best, med = _timed(torch, lambda: gb.valgrad_(v, g, G, Xs, fast=fast), reps)
        par = None
        if not fast:
            ov, og, _, _ = O.evaluate(co, "nan", "cubic", sub(Xs, idx), 3)
            par = bool(np.array_equal(ov, sub(v, idx), equal_nan=True) and np.array_equal(og, sub(g, idx), equal_nan=True))
        rows.append(_row("C2 3-D InterpCubic BCnan 256^3 f64 valgrad, 1e7 queries sorted by 16x16x8-cell block (coherent: the probe hands them to the plain kernel), %s" % ("FAST" if fast else "EXACT"), nq, "points",
                         BYTES_PER_QUERY, best, med, peak, par, traffic.get("c2_sorted_%s_bytes_per_step" % ("fast" if fast else "exact"))))
    # the same valgrad on the RANDOM batch with the packed output layout (GB_OUT_PACKED: one (N+1)-element record per query)
    X = torch.empty((nq, 3), dtype=torch.float64, device=dev)
    gb.synth_uniform(X, SEED_Q, lo=1.0, hi=float(n))
    vg = torch.empty((nq, 4), dtype=torch.float64, device=dev)
    idx = torch.arange(0, nq, 200, device=dev)
    for fast in (False, True):
        best, med = _timed(torch, lambda: gb.valgrad_packed_(vg, G, X, fast=fast), reps)
        par = None
        if not fast:
            ov, og, _, _ = O.evaluate(co, "nan", "cubic", sub(X, idx), 3)
            par = bool(np.array_equal(np.concatenate([ov[:, None], og], axis=1), sub(vg, idx), equal_nan=True))
        rows.append(_row("C2 3-D InterpCubic BCnan 256^3 f64 valgrad, 1e7 uniform queries, packed (N+1) x nq output (new layout: no unpack pass), %s"
                         % ("FAST" if fast else "EXACT"), nq, "points", BYTES_PER_QUERY, best, med, peak, par))
    del G, A, Xs, v, g, X, vg

    # ---- C3: the whole 1e9-query batch, 1e8 at a time; and the prefilter of the 8192^2 image --------------------------
    n, nq, nsl = 8192, 100_000_000, 10
    A = gb.jl_empty((n, n), torch.float32, dev)
    gb.synth_uniform(A, 3)
    G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
    co = G.coefs
    X = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    v = torch.empty(nq, dtype=torch.float32, device=dev)
    g = torch.empty((nq, 2), dtype=torch.float32, device=dev)
    h = torch.empty((nq, 2, 2), dtype=torch.float32, device=dev)
    gb.synth_uniform(X, 14, lo=1.0, hi=float(n))
    gb.valgradhess_(v, g, h, G, X)
    torch.cuda.synchronize()
    tot, par = 0.0, True
    for k in range(nsl):
        gb.synth_uniform(X, 14, lo=1.0, hi=float(n), first=k * nq * 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gb.valgradhess_(v, g, h, G, X)
        e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
        if k in (0, nsl - 1):
            idx = torch.arange(0, nq, 5000, device=dev)
            ov, og, oh, _ = O.evaluate(co, "reflect", "quadratic", sub(X, idx), 7)
            par = par and bool(np.array_equal(ov, sub(v, idx), equal_nan=True) and np.array_equal(og, sub(g, idx), equal_nan=True)
                               and np.array_equal(oh, sub(h, idx), equal_nan=True))
    rows.append(_row("C3 2-D InterpQuadratic BCreflect 8192^2 f32 valgradhess, 1e9 queries (10 slices of 1e8)", nq * nsl, "points", 36, tot, tot, peak, par,
                     traffic.get("c3_plain_blocked_bytes_per_1e8"), note="ms = sum of the 10 slices; traffic is per 1e8-query launch"))
    del X, v, g, h, G
    for fast in (False, True):
        def pf():
            B = A.clone()
            gb.interp_invert_(B, gb.BCreflect, gb.InterpQuadratic, fast=fast)
        try:
            def cl():
                A.clone()
            b0, m0 = _timed(torch, cl, reps)
            best, med = _timed(torch, pf, reps)
        except TypeError:
            continue  # library without the FAST prefilter
        rows.append(_row("C3 prefilter: interp_invert! of the 8192^2 f32 image, both sweeps, %s" % ("FAST (in-line parallel solver)" if fast else "EXACT (LU order)"),
                         n * n, "points", 16, max(best - b0, 1e-6), max(med - m0, 1e-6), peak, None,
                         traffic.get("c3_prefilter_exact_bytes") if not fast else None, note="time of the device copy that precedes the in-place solve subtracted"))
    del A

    # ---- C4 ----------------------------------------------------------------------------------------------------
    n, nq = 64, 100_000_000
    A = gb.jl_empty((n,) * 4, torch.float64, dev)
    gb.synth_uniform(A, 4)
    rng = gb.FloatRange(-63.0, 2.0, 64, 63.0)  # -1.0 : 2/63 : 1.0
    X = torch.empty((nq, 4), dtype=torch.float64, device=dev)
    gb.synth_uniform(X, 15, lo=-1.0 + (-16.0) * 2.0 / 63.0, hi=1.0 + 16.0 * 2.0 / 63.0)  # index coordinates U[1-16, 64+16]
    v = torch.empty(nq, dtype=torch.float64, device=dev)
    idx = torch.arange(0, nq, 5000, device=dev)
    for name, bc in (("periodic", gb.BCperiodic), ("nearest", gb.BCnearest), ("fill", -1.0)):
        G = gb.InterpGrid(A, bc, gb.InterpLinear)
        CG = gb.CoordInterpGrid((rng,) * 4, G)
        best, med = _timed(torch, lambda: gb.getindex_(v, CG, X), reps)
        ov, _, _, _ = O.evaluate(G.coefs, name, "linear", sub(X, idx), 1, affine=[rng.row()] * 4)
        rows.append(_row("C4 4-D InterpLinear BC%s 64^4 f64 getindex via CoordInterpGrid, 1e8 queries" % name, nq, "points", 40, best, med, peak,
                         bool(np.array_equal(ov, sub(v, idx), equal_nan=True)), traffic.get("c4_plain_blocked_%s_bytes_per_1e8" % name)))
        del G, CG
    del A, X, v

    # ---- C5 ----------------------------------------------------------------------------------------------------
    n = 1024
    A = gb.jl_empty((n, n, n), torch.float64, dev)
    gb.synth_uniform(A, 5)
    sizes = [n]
    while sizes[-1] > 17:
        sizes.append(gb.restrict_size(sizes[-1]))

    def down():
        lv = [A]
        for _ in sizes[1:]:
            lv.append(gb.restrict(lv[-1], [True, True, True]))
        return lv

    best_d, med_d = _timed(torch, down, reps)
    levels = down()
    coarse = levels[-1]

    up_sizes = np.array([[s, s, s] for s in reversed(sizes[:-1])]).T  # one column of target sizes per level

    def up():  # prolong(A, sizes) with one column per level: one library call (gb_prolong3_levels)
        return gb.prolong(coarse, up_sizes)

    best_u, med_u = _timed(torch, up, reps)
    k = sizes.index(129)
    a129 = gb.jl_to_numpy(levels[k])
    par = bool(np.array_equal(O.restrict_flags(a129, [True] * 3), gb.jl_to_numpy(levels[k + 1])))
    par = par and bool(np.array_equal(O.prolong_sizes(gb.jl_to_numpy(levels[k + 1]), [129] * 3), gb.jl_to_numpy(gb.prolong(levels[k + 1], [129] * 3))))
    bpu = 8.0 * sum(a ** 3 + b ** 3 for a, b in zip(sizes[:-1], sizes[1:])) / n ** 3
    chain = "->".join(map(str, sizes))
    rows.append(_row("C5 restrict chain %s f64 (fused 3-D marching kernel)" % chain, n ** 3, "points", bpu, best_d, med_d, peak, par, traffic.get("c5_restrict_level0_bytes"),
                     note="unit = level-0 point; traffic = the 1024^3 -> 513^3 launch alone (9.66 GB algorithmic); parity on the 129^3 level (full-size levels: tests/test_gpu_fullsize.py)"))
    rows.append(_row("C5 prolong chain back up (fused 3-D marching kernel)", n ** 3, "points", bpu, best_u, med_u, peak, par, traffic.get("c5_prolong_level0_bytes"), note="traffic = the 513^3 -> 1024^3 launch alone"))
    return rows


def c3_sharded(gb, torch, dist, np, O, peak, rank, world, dev):
    """C3's 1e9 queries split over the ranks (strong scaling): contiguous ranges, coefficients broadcast once."""
    import grid_jl_b200.dist as D

    n, total, slice_q = 8192, 1_000_000_000, 50_000_000
    G = None
    if rank == 0:
        A = gb.jl_empty((n, n), torch.float32, dev)
        gb.synth_uniform(A, 3)
        G = gb.InterpGrid(A, gb.BCreflect, gb.InterpQuadratic)
        del A
    D.default_comm()  # (NCCL communicator of the library, created once per process: ~0.5 s, not part of the broadcast)
    dist.barrier()
    t0 = time.perf_counter()
    G = D.broadcast_grid(G, gb.BCreflect, gb.InterpQuadratic, src=0)
    torch.cuda.synchronize()
    t_bcast = time.perf_counter() - t0
    q0, q1 = D.shard_range(total, rank, world)
    X = torch.empty((slice_q, 2), dtype=torch.float32, device=dev)
    v = torch.empty(slice_q, dtype=torch.float32, device=dev)
    g = torch.empty((slice_q, 2), dtype=torch.float32, device=dev)
    h = torch.empty((slice_q, 2, 2), dtype=torch.float32, device=dev)
    gb.synth_uniform(X, 14, lo=1.0, hi=float(n), first=q0 * 2)
    gb.valgradhess_(v, g, h, G, X)  # warm-up
    torch.cuda.synchronize()
    dist.barrier()
    ms, par = 0.0, True
    for s0 in range(q0, q1, slice_q):
        cnt = min(slice_q, q1 - s0)
        gb.synth_uniform(X[:cnt], 14, lo=1.0, hi=float(n), first=s0 * 2)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        gb.valgradhess_(v[:cnt], g[:cnt], h[:cnt], G, X[:cnt])
        e1.record()
        torch.cuda.synchronize()
        ms += e0.elapsed_time(e1)
        if s0 == q0 and rank in (0, world - 1):
            idx = torch.arange(0, cnt, 10_000, device=dev)
            ov, og, oh, _ = O.evaluate(G.coefs, "reflect", "quadratic", X[idx].cpu().numpy(), 7)
            par = bool(np.array_equal(ov, v[idx].cpu().numpy(), equal_nan=True) and np.array_equal(oh, h[idx].cpu().numpy(), equal_nan=True))
    t = torch.tensor([ms, 0.0 if par else 1.0], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t[0].item())
    val = total / (ms_max * 1e-3) / 1e9
    return {"workload": "C3 2-D InterpQuadratic BCreflect 8192^2 f32 valgradhess, 1e9 queries sharded over %d GPUs" % world, "scaling": "strong",
            "ms": ms_max, "value": val, "unit": "Gpoints/s", "queries_per_gpu": q1 - q0, "broadcast_ms": t_bcast * 1e3,
            "roofline": {"bound": "hbm", "achieved": val * 36 / world, "peak": peak, "unit": "GB/s per GPU", "frac": val * 36 / world / peak},
            "parity_vs_oracle": float(t[1].item()) == 0.0}


def c5_slab(gb, torch, dist, np, O, peak, rank, world, dev, reps=3):
    """The C5 chain with the arrays slab-sharded along dim 3; one halo exchange per level (NCCL)."""
    import grid_jl_b200.dist as D

    n = 1024
    sizes = [n]
    while sizes[-1] > 17:
        sizes.append(gb.restrict_size(sizes[-1]))
    f0, f1 = D.shard_range(n, rank, world)
    local = gb.jl_empty((n, n, f1 - f0), torch.float64, dev)
    gb.synth_uniform(local, 5, first=f0 * n * n)  # this rank's planes of the same synthetic array
    res = {"workload": "C5 restrict / prolong chain %s f64, slab-sharded along dim 3 over %d GPUs" % ("->".join(map(str, sizes)), world), "scaling": "strong"}
    out = {}
    for direction in ("restrict", "prolong"):
        best = None
        for rep in range(reps + 1):
            stats = []
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if direction == "restrict":
                cur, lv = local, [local]
                for L in sizes[:-1]:
                    cur = D.slab_restrict3(cur, L, rank, world, stats=stats)
                    lv.append(cur)
            else:
                cur = coarse
                for c, s in zip(reversed(sizes[1:]), reversed(sizes[:-1])):
                    cur = D.slab_prolong3(cur, c, (s, s, s), rank, world, stats=stats)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if rep > 0 and (best is None or ms < best[0]):
                best = (ms, stats)
        if direction == "restrict":
            levels, coarse = lv, lv[-1]
        t = torch.tensor([best[0]], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_levels = float(t.item())
        # the same chain as ONE library call -- the reference's restrict(A, flags) / prolong(A, sizes) with one column per level
        # (gb_restrict3_sharded_levels / gb_prolong3_sharded_levels): no per-level return to the host, no per-level statistics
        comm = D.default_comm()
        best1, res1 = None, None
        for rep in range(reps + 1):
            torch.cuda.synchronize()
            dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if direction == "restrict":
                res1 = comm.restrict3_levels([local], (n, n, n), len(sizes) - 1, 1.0)[0]
            else:
                res1 = comm.prolong3_levels([coarse], (sizes[-1],) * 3, [(s,) * 3 for s in reversed(sizes[:-1])])[0]
            e1.record()
            torch.cuda.synchronize()
            if rep > 0:
                best1 = e0.elapsed_time(e1) if best1 is None else min(best1, e0.elapsed_time(e1))
        same_bits = bool(torch.equal(res1, cur))
        t = torch.tensor([best1, 0.0 if same_bits else 1.0], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t[0].item())
        one_call_ok = float(t[1].item()) == 0.0
        bpu = 8.0 * sum(a ** 3 + b ** 3 for a, b in zip(sizes[:-1], sizes[1:])) / n ** 3
        val = n ** 3 / (ms * 1e-3) / 1e9
        out[direction] = {"ms": ms, "value": val, "unit": "Gpoints/s (level-0 points)",
                          "roofline": {"bound": "hbm", "achieved": val * bpu / world, "peak": peak, "unit": "GB/s per GPU", "frac": val * bpu / world / peak},
                          "call": "one library call per chain (restrict(A, flags) / prolong(A, sizes) with one column per level)",
                          "one_call_equals_level_calls_bitwise": one_call_ok,
                          "ms_level_calls": ms_levels, "levels_rank0": best[1]}
    # parity: the 129^3 level gathered on rank 0 against the oracle applied to the gathered 257^3 level
    k = sizes.index(257)
    parts = [None] * world
    dist.all_gather_object(parts, (gb.jl_to_numpy(levels[k]), gb.jl_to_numpy(levels[k + 1])))
    par = None
    if rank == 0:
        fine = np.concatenate([p[0] for p in parts], axis=2)
        coarse_ref = O.restrict_flags(np.asfortranarray(fine), [True] * 3)
        par = bool(np.array_equal(coarse_ref, np.concatenate([p[1] for p in parts], axis=2)))
    res.update(out)
    res["parity_vs_oracle"] = par
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fast", action="store_true", help="time the GB_FAST (separable FMA) contraction instead of the bit-exact default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip extra.configs / extra.c3_sharded / extra.c5_slab")
    args = ap.parse_args()
    # stdout carries exactly ONE JSON line: everything else that writes to fd 1 (NCCL's version banner, library
    # chatter) is sent to stderr while the benchmark runs; the line itself goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)

    import numpy as np
    import torch
    import torch.distributed as dist

    import grid_jl_b200 as gb

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = N_GRID
    peak, how = measured_peak()
    traffic = load_traffic()

    # ---- grid construction: synth samples on device, pad + prefilter on rank 0, broadcast coefs ----
    pref_ms = None
    if rank == 0:
        A = gb.jl_empty((n, n, n), torch.float64, dev)
        gb.synth_uniform(A, SEED_GRID)
        G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)  # warm-up (allocations, module load)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        times = []
        for _ in range(3):
            del G
            e0.record()
            G = gb.InterpGrid(A, gb.BCnan, gb.InterpCubic)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        pref_ms = min(times)
        del A
    if world > 1:
        stored = (n + 2,) * 3
        buf = G.coefs_device().clone() if rank == 0 else gb.jl_empty(stored, torch.float64, dev)
        flat = buf.permute(2, 1, 0).reshape(-1)  # the underlying contiguous storage
        dist.broadcast(flat, src=0)
        if rank != 0:
            G = gb.InterpGrid.from_coefs(buf, gb.BCnan, gb.InterpCubic)
        del buf, flat
    torch.cuda.synchronize()

    # ---- device-resident step -------------------------------------------------------------------
    X = torch.empty((NQ, 3), dtype=torch.float64, device=dev)
    gb.synth_uniform(X, SEED_Q + 1000 * rank, lo=1.0, hi=float(n))
    val = torch.empty(NQ, dtype=torch.float64, device=dev)
    grad = torch.empty((NQ, 3), dtype=torch.float64, device=dev)

    def step():
        gb.valgrad_(val, grad, G, X, fast=args.fast)

    clk = ClockSampler(local)
    clk.__enter__()  # samples SM clock / throttle reasons from the warm-up until the GPU work below is done
    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    l0 = gb.launch_count()
    t_win0 = time.perf_counter()
    ev[0].record()
    for i in range(args.steps):
        step()
        ev[i + 1].record()
    torch.cuda.synchronize()
    t_win1 = time.perf_counter()
    launches = gb.launch_count() - l0
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = ev[0].elapsed_time(ev[-1])
    per = sorted(ev[i].elapsed_time(ev[i + 1]) for i in range(args.steps))
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms_max = float(t.item())
    ms_per_step = total_ms_max / args.steps
    value = world * NQ / (ms_per_step * 1e-3) / 1e9

    # ---- stage timing of the step (CUDA events on the launch stream, recorded inside the library) ----------
    def stage_profile(fast, reps=5):
        gb.profile_enable(True)
        acc = {}
        for _ in range(reps):
            gb.valgrad_(val, grad, G, X, fast=fast)
            for k, v in gb.profile_read().items():
                acc[k] = acc.get(k, 0.0) + v / reps
        gb.profile_enable(False)
        return acc

    stages = stage_profile(args.fast)
    # the other contraction, for the record (same step, same timing method, fewer steps)
    other = None
    if rank == 0 and world == 1:
        of = not args.fast
        for _ in range(3):
            gb.valgrad_(val, grad, G, X, fast=of)
        torch.cuda.synchronize()
        oe = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        osteps = max(3, min(args.steps, 10))
        oe[0].record()
        for _ in range(osteps):
            gb.valgrad_(val, grad, G, X, fast=of)
        oe[1].record()
        torch.cuda.synchronize()
        oms = oe[0].elapsed_time(oe[1]) / osteps
        ostages = stage_profile(of)
        okern = ostages.get("eval_brick", oms)
        other = {"contraction": "FAST" if of else "EXACT", "value": NQ / (oms * 1e-3) / 1e9, "unit": "Gpoints/s", "ms_per_step": oms,
                 "stages_ms": ostages, "roofline_frac_kernel": BYTES_PER_QUERY * NQ / (okern * 1e-3) / 1e9 / peak,
                 "roofline_frac_step": BYTES_PER_QUERY * NQ / (oms * 1e-3) / 1e9 / peak}
        gb.valgrad_(val, grad, G, X, fast=args.fast)  # leave the headline mode's results in val / grad
        torch.cuda.synchronize()

    # ---- end to end through the host-buffer C-ABI call ----------------------------------------------
    e2e = e2e_pageable = None
    if not args.no_e2e:
        def host_run(Xn, vn, gn, ksteps):
            for _ in range(2):
                gb.valgrad_(vn, gn, G, Xn, fast=args.fast)
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            for _ in range(ksteps):
                gb.valgrad_(vn, gn, G, Xn, fast=args.fast)  # returns when the results are in vn/gn
            dt = (time.perf_counter() - t0) / ksteps
            te = torch.tensor([dt], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return float(te.item())

        Xh = torch.empty((NQ, 3), dtype=torch.float64).pin_memory()
        Xh.copy_(X)
        vh = torch.empty(NQ, dtype=torch.float64).pin_memory()
        gh = torch.empty((NQ, 3), dtype=torch.float64).pin_memory()
        ksteps = max(3, min(args.steps, 10))
        dt = host_run(Xh.numpy(), vh.numpy(), gh.numpy(), ksteps)
        assert torch.equal(vh.to(dev), val) or args.fast, "host-path results differ from the device path"
        e2e = {"value": world * NQ / dt / 1e9, "unit": "Gpoints/s", "h2d_bytes_per_step": NQ * 24, "d2h_bytes_per_step": NQ * 32,
               "ms_per_step": dt * 1e3, "steps": ksteps, "host_memory": "pinned"}
        # the same call with pageable buffers: what a Julia Array (plain malloc) gives unless the caller pins it
        Xp, vp, gp = np.array(Xh.numpy(), copy=True), np.empty(NQ), np.empty((NQ, 3))
        del Xh, vh, gh
        dtp = host_run(Xp, vp, gp, max(2, ksteps // 2))
        assert np.array_equal(vp, val.cpu().numpy()) or args.fast
        e2e_pageable = {"value": world * NQ / dtp / 1e9, "unit": "Gpoints/s", "ms_per_step": dtp * 1e3, "host_memory": "pageable (numpy / malloc)",
                        "staging": "library-owned pinned slots, %s host threads (GB_HOST_THREADS)" % os.environ.get("GB_HOST_THREADS", "min(8, cores/2)"),
                        "h2d_bytes_per_step": NQ * 24, "d2h_bytes_per_step": NQ * 32}
        del Xp, vp, gp

    clk.__exit__()

    # ---- the other BASELINE configs ---------------------------------------------------------------------------
    from oracle import oracle as O  # the checker (parity flags of the extras, cpu_baseline below)

    extra_rows = sharded = slab = None
    extra_err = None
    if not args.no_extra:
        try:
            if world == 1:
                torch.cuda.empty_cache()
                extra_rows = extra_configs(gb, torch, np, O, peak, traffic)
            else:
                torch.cuda.empty_cache()
                sharded = c3_sharded(gb, torch, dist, np, O, peak, rank, world, dev)
                torch.cuda.empty_cache()
                slab = c5_slab(gb, torch, dist, np, O, peak, rank, world, dev)
        except Exception as e:  # the headline line must survive a failure in the extras; the error is reported in it
            import traceback

            extra_err = "%s: %s | %s" % (type(e).__name__, e, traceback.format_exc().strip().splitlines()[-3:])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel of the step ------------------------------------------------------
    step_ms = total_ms / args.steps
    kernel_ms = stages.get("eval_brick", step_ms)
    achieved = BYTES_PER_QUERY * NQ / (kernel_ms * 1e-3) / 1e9
    step_achieved = BYTES_PER_QUERY * NQ / (step_ms * 1e-3) / 1e9
    tkey = "fast" if args.fast else "exact"
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic.get("eval_brick_%s_bytes_per_launch" % tkey),
                "peak_source": how, "kernel": "eval_brick_kernel<double,3,Cubic,BCF_NIL,valgrad,%s>" % ("FAST" if args.fast else "EXACT"),
                "algorithmic_bytes_per_launch": BYTES_PER_QUERY * NQ, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
                "stages_ms": stages, "step_achieved": step_achieved, "step_frac": step_achieved / peak,
                "step_traffic": traffic.get("c2_step_%s_bytes" % tkey), "step_traffic_by_stage": traffic.get("c2_step_%s_by_stage" % tkey),
                "traffic_source": traffic.get("source"),
                "frac_of_nominal_8TBs": achieved / 8000.0,
                # what actually bounds the EXACT kernel: the FP64 pipe (64 lanes/SM, one mul or add per lane per clock;
                # ncu sm__pipe_fp64_cycles_active 80 %).  816 unfusable mul/add of the
                # reference's contraction + ~130 for weights / coordinates per query; FAST: 192 FMA + ~75.
                "fp64_pipe": (lambda ops, clk_: {"ops_per_query": ops, "achieved_Tops": ops * NQ / (kernel_ms * 1e-3) / 1e12,
                                                 "peak_Tops": 64 * 148 * clk_ * 1e6 / 1e12,
                                                 "frac": ops * NQ / (kernel_ms * 1e-3) / (64 * 148 * clk_ * 1e6)})(
                                                     267 if args.fast else 945, (clk.max_mhz or 1965)),
                "limiter": "FP64 pipe (816 dependent mul/add per query in the reference's order)" if not args.fast else "shared-memory gathers + FP64 pipe"}

    # ---- CPU baseline (oracle port) on a bounded sample ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = int(os.environ.get("GB_REF_SAMPLE", str(NQ)))  # the whole batch: ~2 s on 16 cores (+ ~1 s single-core subsample)
        co = G.coefs
        Xs = X[:sample].cpu().numpy()
        threads = host_cores()
        O.evaluate(co, "nan", "cubic", Xs[: sample // 10], 3, nthreads=threads)
        t0 = time.perf_counter()
        ov, og, _, _ = O.evaluate(co, "nan", "cubic", Xs, 3, nthreads=threads)
        dt_all = time.perf_counter() - t0
        one = max(1, sample // 20)
        t0 = time.perf_counter()
        O.evaluate(co, "nan", "cubic", Xs[:one], 3, nthreads=1)
        dt_one = time.perf_counter() - t0
        parity = bool(np.array_equal(ov, val[:sample].cpu().numpy(), equal_nan=True) and np.array_equal(og, grad[:sample].cpu().numpy(), equal_nan=True))
        cpu = {"value": sample / dt_all / 1e9, "unit": "Gpoints/s", "cores": threads, "kind": "port",
               "sample": f"first {sample} of the 1e7 queries, OpenMP over queries", "single_core_value": one / dt_one / 1e9,
               "bit_identical_to_gpu": parity if not args.fast else None}

    line = {
        "metric": METRIC, "value": value, "unit": "Gpoints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": c2_config(world, "FAST" if args.fast else "EXACT"),
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary((t_win0, t_win1)),
        "extra": {"prefilter_ms": pref_ms, "prefilter_note": "InterpGrid(A,BCnan,InterpCubic) on device: pad + 3 prefilter sweeps of 258^3 f64",
                  "prefilter_roofline_frac": (64.0 * (n + 2) ** 3 / (pref_ms * 1e-3) / 1e9 / peak) if pref_ms else None,
                  "step_ms_min": per[0], "step_ms_median": per[len(per) // 2], "step_ms_max": per[-1],
                  "other_contraction": other, "e2e_pageable": e2e_pageable,
                  "with_prefilter_ms_per_step": (ms_per_step + pref_ms) if pref_ms else None,
                  "configs": extra_rows, "c3_sharded": sharded, "c5_slab": slab, "extra_error": extra_err},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
