#!/usr/bin/env python
"""bench.py -- Gpoints/s of InterpGrid valgrad on B200 (BASELINE.json metric), one JSON line.

Workload (config.workload = "C2"): BASELINE configs[1] -- 3-D InterpCubic BCnan on a 256^3 Float64
grid (stored 258^3 after the cubic padding) incl. the interp_invert! prefilter at construction,
1e7 uniform-random queries per GPU, valgrad (value + 3 gradient components).

A "step" is one pass of valgrad over one batch of 1e7 queries.
  value  device-resident: queries / outputs already in HBM; CUDA-event timed on the launch stream.
  e2e    the same step through the C-ABI with HOST buffers (pinned): H2D of the queries and D2H of
         value+gradient inside the timed region (gb_eval, mem=GB_HOST).
  roofline  the dominant kernel of the step is eval_brick_kernel<double,3,Cubic,nil,valgrad,...>; its
         average duration is measured live with CUDA events on the launch stream (gb_profile_*),
         algorithmic bytes = 56 B/query (24 in + 32 out) x 1e7 queries per launch ->
         achieved GB/s vs the measured HBM peak.  roofline.step_* gives the same figure for the whole
         step (binning sort + brick evaluation + un-permute), which is what `value` measures.
  contraction  the headline runs the library default, EXACT (the reference's arithmetic order,
         bit-identical to the oracle); extra.fast repeats the step with GB_FAST (separable FMA,
         within north_star's 4 ulp of the magnitude scale -- tests/test_gpu_parity.py).
  cpu_baseline  the C++ restatement of the reference algorithm (oracle, "port") on the host cores,
         bounded sample.
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


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


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
    A = O.synth(np.float64, n ** 3, SEED_GRID).reshape((n, n, n), order="F")
    t0 = time.perf_counter()
    co = O.make_coefs(A, "nan", "cubic")
    t_pref = time.perf_counter() - t0
    threads = O.max_threads()
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
        "impl": "reference", "metric": "valgrad throughput, 3-D InterpCubic BCnan 256^3 Float64", "value": val, "unit": "Gpoints/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2", "grid": "256^3 f64 (stored 258^3)", "order": "InterpCubic", "bc": "BCnan", "outputs": "valgrad",
                   "queries_per_step": sample, "note": "Julia cannot run in this image, "
                   "this is the C++ restatement of the reference algorithm (oracle)"},
        "cpu_baseline": {"value": val, "unit": "Gpoints/s", "cores": threads, "kind": "port", "sample": f"{sample} of 1e7 queries per step, OpenMP over queries",
                         "prefilter_s": t_pref},
        "e2e": {"value": val, "unit": "Gpoints/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--fast", action="store_true", help="time the GB_FAST (separable FMA) contraction instead of the bit-exact default")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
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
        other = {"contraction": "FAST" if of else "EXACT", "value": NQ / (oms * 1e-3) / 1e9, "unit": "Gpoints/s", "ms_per_step": oms,
                 "stages_ms": stage_profile(of)}
        gb.valgrad_(val, grad, G, X, fast=args.fast)  # leave the headline mode's results in val / grad
        torch.cuda.synchronize()

    # ---- end to end through the host-buffer C-ABI call ----------------------------------------------
    e2e = None
    if not args.no_e2e:
        Xh = torch.empty((NQ, 3), dtype=torch.float64).pin_memory()
        Xh.copy_(X)
        vh = torch.empty(NQ, dtype=torch.float64).pin_memory()
        gh = torch.empty((NQ, 3), dtype=torch.float64).pin_memory()
        Xn, vn, gn = Xh.numpy(), vh.numpy(), gh.numpy()
        ksteps = max(3, min(args.steps, 10))
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
        dt = float(te.item())
        assert torch.equal(vh.to(dev), val) or args.fast, "host-path results differ from the device path"
        e2e = {"value": world * NQ / dt / 1e9, "unit": "Gpoints/s", "h2d_bytes_per_step": NQ * 24, "d2h_bytes_per_step": NQ * 32,
               "ms_per_step": dt * 1e3, "steps": ksteps, "host_memory": "pinned"}

    clk.__exit__()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel of the step ------------------------------------------------------
    peak, how = measured_peak()
    step_ms = total_ms / args.steps
    kernel_ms = stages.get("eval_brick", step_ms)
    achieved = BYTES_PER_QUERY * NQ / (kernel_ms * 1e-3) / 1e9
    step_achieved = BYTES_PER_QUERY * NQ / (step_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("eval_brick_fast_bytes_per_launch" if args.fast else "eval_brick_exact_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "peak_source": how, "kernel": "eval_brick_kernel<double,3,Cubic,BCF_NIL,valgrad,%s>" % ("FAST" if args.fast else "EXACT"),
                "algorithmic_bytes_per_launch": BYTES_PER_QUERY * NQ, "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
                "stages_ms": stages, "step_achieved": step_achieved, "step_frac": step_achieved / peak,
                "frac_of_nominal_8TBs": achieved / 8000.0,
                # what actually bounds the EXACT kernel: the FP64 pipe (64 lanes/SM, one mul or add per lane per clock;
                # ncu sm__pipe_fp64_cycles_active 80 %, profiles/r1_v12_exact_full.txt).  816 unfusable mul/add of the
                # reference's contraction + ~130 for weights / coordinates per query; FAST: 192 FMA + ~75.
                "fp64_pipe": (lambda ops, clk: {"ops_per_query": ops, "achieved_Tops": ops * NQ / (kernel_ms * 1e-3) / 1e12,
                                                "peak_Tops": 64 * 148 * clk * 1e6 / 1e12,
                                                "frac": ops * NQ / (kernel_ms * 1e-3) / (64 * 148 * clk * 1e6)})(
                                                    267 if args.fast else 945, (clk.max_mhz or 1965)),
                "limiter": "FP64 pipe (816 dependent mul/add per query in the reference's order)" if not args.fast else "shared-memory gathers + FP64 pipe"}

    # ---- CPU baseline (oracle port) on a bounded sample ---------------------------------------------------
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle as O

        sample = int(os.environ.get("GB_REF_SAMPLE", str(NQ)))  # the whole batch: ~2 s on 16 cores (+ ~1 s single-core subsample)
        co = G.coefs
        Xs = X[:sample].cpu().numpy()
        threads = O.max_threads()
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
        "metric": "valgrad throughput, 3-D InterpCubic BCnan 256^3 Float64", "value": value, "unit": "Gpoints/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2", "grid": "256^3 f64 (stored 258^3)", "order": "InterpCubic", "bc": "BCnan", "outputs": "valgrad",
                   "queries_per_gpu_per_step": NQ, "query_distribution": "uniform [1,256]^3, splitmix64", "contraction": "FAST" if args.fast else "EXACT",
                   "l2": "inputs larger than L2 (240 MB in + 320 MB out + 137 MB grid per step vs 126 MB L2)",
                   "parallelism": f"query-sharded x{world}, coefficients replicated by NCCL broadcast"},
        "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clk.summary((t_win0, t_win1)),
        "extra": {"prefilter_ms": pref_ms, "prefilter_note": "InterpGrid(A,BCnan,InterpCubic) on device: pad + 3 prefilter sweeps of 258^3 f64",
                  "step_ms_min": per[0], "step_ms_median": per[len(per) // 2], "step_ms_max": per[-1],
                  "other_contraction": other,
                  "with_prefilter_ms_per_step": (ms_per_step + pref_ms) if pref_ms else None},
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
