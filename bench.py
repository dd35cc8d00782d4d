#!/usr/bin/env python
"""Benchmark of the weather-sim time-stepping hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
                    [--workload NAME] [--kernel auto|stage_direct|step_fused_reg|step_fused_tma]

Metric (BASELINE.json): grid cell-updates/s of the RK4 Shallow-Water step. A "step" is one full time step
over the whole grid. The default workload is BASELINE config 2 (SWE 8192x8192 fp32 RK4) at N = 1 and weak
scaling of it at N > 1 (the same 8192x8192 row slab per GPU, global grid 8192 x 8192*N, NCCL ghost-row
exchange inside libweather_b200.so). One process per GPU (torchrun); torch.distributed (NCCL) is used only
for the barrier and the max-over-ranks reduction of the timings -- the product path never imports torch.

Other BASELINE configs are available as --workload (their lines carry their own metric name):
    swe8192_euler   SWE 8192^2 fp32 Euler            (1 fused pass, 24 B/cell-step: the HBM-bound case)
    baro16384_f64   Barotropic 16384^2 fp64 "RK4"    (reference semantics: SWE tendencies + RK2)
    prim2048x64     Primitive 2048^2 x 64 levels fp32 RK2 (+ constant T/p drift)
    swe32768_rk4    SWE 32768^2 fp32 RK4, STRONG scaling: the global grid is split over the N GPUs

--impl reference times the reference's own CPU implementation (oracle/_ref/libws_ref.so, the patched
reference build; the oracle port if that is absent) on the host cores, on a bounded band of the workload.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(ROOT, "nvidia-jetson-workload_b200")]

UNIT = "cell-updates/s"
# bpc_stage / bpc_step: algorithmic bytes per cell per step of a one-pass-per-RK-stage
# design and of the whole-step kernels (read y_n once, write y_n+1 once; + T,p read/write for Primitive).
WORKLOADS = {
    # schema/self-test size (tests/test_bench_gpu.py), not a benchmark
    "tiny_swe512_rk4": dict(W=512, rows=512, model="shallow_water", integ="rk4", dtype="f32", levels=1,
                            scaling="weak", metric="grid cell-updates/sec (RK4 SWE step)",
                            desc="SWE {W}x{H} fp32 RK4 (self-test size)", bpc_stage=168, bpc_step=24),
    "swe8192_rk4": dict(W=8192, rows=8192, model="shallow_water", integ="rk4", dtype="f32", levels=1,
                        scaling="weak", metric="grid cell-updates/sec (RK4 SWE step)",
                        desc="SWE {W}x{H} fp32 RK4 (reference-parity combine)", bpc_stage=168, bpc_step=24),
    "swe8192_euler": dict(W=8192, rows=8192, model="shallow_water", integ="euler", dtype="f32", levels=1,
                          scaling="weak", metric="grid cell-updates/sec (Euler SWE step)",
                          desc="SWE {W}x{H} fp32 Euler", bpc_stage=24, bpc_step=24),
    "baro16384_f64": dict(W=16384, rows=16384, model="barotropic", integ="rk4", dtype="f64", levels=1,
                          scaling="weak", metric="grid cell-updates/sec (Barotropic fp64 RK4->RK2 step)",
                          desc="Barotropic {W}x{H} fp64, RK4 requested = SWE tendencies + RK2 (reference semantics)",
                          bpc_stage=120, bpc_step=48),
    "prim2048x64": dict(W=2048, rows=2048, model="primitive", integ="rk2", dtype="f32", levels=64,
                        scaling="weak", metric="grid cell-updates/sec (Primitive 64-level RK2 step)",
                        desc="Primitive {W}x{H}x64 levels fp32 RK2 (SWE tendencies per level + constant T/p drift)",
                        bpc_stage=100, bpc_step=40),
    # true IEEE division: 2dx, 2dy are not powers of two, so (a-b)/(2dx) cannot become a multiplication
    "swe8192_rk4_div": dict(W=8192, rows=8192, model="shallow_water", integ="rk4", dtype="f32", levels=1,
                            scaling="weak", metric="grid cell-updates/sec (RK4 SWE step)", dx=0.8, dy=1.7,
                            desc="SWE {W}x{H} fp32 RK4, dx=0.8 dy=1.7 (true-division kernels)", bpc_stage=168,
                            bpc_step=24),
    # NOT a reference configuration: the extended physics opt-in (beta plane + viscosity + diffusivity), textbook RK4
    "swe8192_rk4_ext": dict(W=8192, rows=8192, model="shallow_water", integ="rk4", dtype="f32", levels=1,
                            scaling="weak", metric="grid cell-updates/sec (RK4 step, extended physics)",
                            extended=(1.0e-5, 0.05, 0.02), classical=True,
                            desc="beta-plane shallow water with viscosity/diffusivity {W}x{H} fp32, textbook RK4 "
                                 "(WSB_PHYSICS_EXTENDED: not in the reference)", bpc_stage=192, bpc_step=24),
    # NOT a reference configuration: extended physics on the Primitive model (p, T, q transported by the flow), per-stage
    # kernels; algorithmic bytes per cell-step of RK2: (6 + 9) fields for u, v, h + (8 + 11) for the tracer stages
    "prim2048x64_ext": dict(W=2048, rows=2048, model="primitive", integ="rk2", dtype="f32", levels=64,
                            scaling="weak", metric="grid cell-updates/sec (Primitive 64-level RK2 step, extended physics)",
                            extended=(1.0e-5, 0.05, 0.02),
                            desc="Primitive {W}x{H}x64 levels fp32 RK2, beta plane + viscosity + tracer transport of "
                                 "p, T, q (WSB_PHYSICS_EXTENDED: not in the reference)", bpc_stage=136, bpc_step=136),
    "swe32768_rk4": dict(W=32768, rows=32768, model="shallow_water", integ="rk4", dtype="f32", levels=1,
                         scaling="strong", metric="grid cell-updates/sec (RK4 SWE step)",
                         desc="SWE {W}x{H} fp32 RK4 (reference-parity combine), strong scaling",
                         bpc_stage=168, bpc_step=24),
}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU through NVML while the timed region runs."""

    def __init__(self, index, period_s=0.005):
        self.index, self.period = index, period_s
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4),
            "hw_power_brake": getattr(nv, "nvmlClocksEventReasonHwPowerBrakeSlowdown", 0x80),
        }
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or getattr(
            nv, "nvmlDeviceGetCurrentClocksThrottleReasons")
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = get_reasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(self.period)

    def __enter__(self):
        if self.nv is not None:
            self._thr = threading.Thread(target=self._loop, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        if self._thr is not None:
            self._thr.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


INTEG_CODE = {"euler": 0, "rk2": 1, "rk4": 2}
MODEL_CODE = {"shallow_water": 0, "barotropic": 1, "primitive": 2}


def fill_bump_slab(h, r0, W, H_global):
    """Gaussian height bump centred on the GLOBAL grid (IC-A), written row-block wise."""
    sigma = 0.1 * min(W, H_global)
    xx = (np.arange(W, dtype=np.float64)[None, :] - (W - 1) / 2.0) ** 2
    nrows = h.shape[-2]
    for b in range(0, nrows, 512):
        e = min(b + 512, nrows)
        yy = (np.arange(r0 + b, r0 + e, dtype=np.float64)[:, None] - (H_global - 1) / 2.0) ** 2
        h[..., b:e, :] = (10.0 + np.exp(-(xx + yy) / (2.0 * sigma * sigma))).astype(h.dtype)


def fill_bump_fast(h, r0, W, H_global):
    """The same bump as an outer product exp(-x^2/2s^2) * exp(-y^2/2s^2) (one multiply per cell instead of one exp):
    input data for the sub-lines, which time kernels and compare nothing."""
    sigma = 0.1 * min(W, H_global)
    ex = np.exp(-((np.arange(W, dtype=np.float64) - (W - 1) / 2.0) ** 2) / (2.0 * sigma * sigma)).astype(h.dtype)
    nrows = h.shape[-2]
    ey = np.exp(-((np.arange(r0, r0 + nrows, dtype=np.float64) - (H_global - 1) / 2.0) ** 2) /
                (2.0 * sigma * sigma)).astype(h.dtype)
    for b in range(0, nrows, 1024):
        e = min(b + 1024, nrows)
        np.multiply(ey[b:e, None], ex[None, :], out=h[..., b:e, :])
        h[..., b:e, :] += h.dtype.type(10.0)


def source_stamp():
    """git blob hashes of the kernel sources: profiles/*.json derived from an ncu capture carry the stamp of the
    sources they were captured from, and the bench line says whether they still match."""
    import hashlib
    out = {}
    for f in ("wsb_step_tma.cu", "wsb_arith.cuh"):
        data = open(os.path.join(ROOT, "nvidia-jetson-workload_b200", "csrc", f), "rb").read()
        out[f] = hashlib.sha1(b"blob %d\0" % len(data) + data).hexdigest()[:12]
    return out


# ------------------------------------------------------------------------------ reference arm --
def cpu_model():
    try:
        with open("/proc/cpuinfo") as f:
            for ln in f:
                if ln.startswith("model name"):
                    return ln.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def time_reference_cpu(width, rows, steps, warmup, model=0, integ=2, threads=None):
    """Times the reference's CPU implementation (weather_simulation.cpp:117-158) on a (rows x width) band."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle_py

    # torchrun exports OMP_NUM_THREADS=1 to every rank, and libgomp may already be loaded: the team size is set through
    # the OpenMP runtime itself and the number reported is the one the reference's parallel loop really gets
    want = threads or host_threads()
    os.environ["OMP_NUM_THREADS"] = str(want)
    cores = (oracle_py.reference_omp_threads(want) if oracle_py.reference_available()
             else oracle_py.oracle_omp_threads(want))
    h = np.empty((rows, width), np.float32)
    fill_bump_slab(h, 0, width, rows)
    u = np.zeros_like(h)
    v = np.zeros_like(h)
    if oracle_py.reference_available():
        kind = "reference"
        sim = oracle_py.Reference(width, rows, model, integ, coriolis_f=0.1)
    else:
        kind = "port"
        oracle_py.build_oracle()
        sim = oracle_py.Oracle(width, rows, model, integ, coriolis_f=0.1)
    sim.set_state(u, v, h)
    if warmup:
        sim.step(warmup)
    t0 = time.perf_counter()
    sim.step(steps)
    dt = time.perf_counter() - t0
    sim.close()
    return {"value": width * rows * steps / dt, "seconds": dt, "kind": kind, "cores": cores,
            "ms_per_step": dt / steps * 1e3}


def config1_reference_cpu():
    """BASELINE config 1 exactly as specified (SWE 256^2 fp32 Euler, 1000 steps) on the reference's CPU path
    ("CPU baseline timing")."""
    r = time_reference_cpu(256, 256, 1000, 0, MODEL_CODE["shallow_water"], INTEG_CODE["euler"])
    return {"value": r["value"], "unit": UNIT, "seconds": r["seconds"], "cores": r["cores"], "kind": r["kind"],
            "workload": "SWE 256x256 fp32 Euler, 1000 steps (BASELINE config 1)"}


def run_reference_arm(args, wl, rank):
    if rank != 0:
        return
    W = wl["W"]
    # a band of the workload sized so that K + W steps finish in ~100 s at the reference's ~8 Mcell/s (RK4)
    rate = 8.0e6 if wl["integ"] != "euler" else 40.0e6
    cells = 100.0 * rate / max(args.steps + args.warmup, 1)
    rows = int(min(wl["rows"], max(64, cells // W)))
    r = time_reference_cpu(W, rows, args.steps, args.warmup, MODEL_CODE[wl["model"]], INTEG_CODE[wl["integ"]])
    note = "" if wl["dtype"] == "f32" and wl["levels"] == 1 else \
        " (the reference computes in fp32 on a single 2-D level whatever the configuration asks, SURVEY.md F7-F8)"
    sample = (f"{r['kind']} CPU path (oracle/_ref, OpenMP team of {r['cores']} threads, measured) on a {W}x{rows} band of the "
              f"workload, {args.steps} steps after {args.warmup} warm-up{note}")
    line = {
        "impl": "reference", "metric": wl["metric"], "value": r["value"], "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
        "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": wl["desc"].format(W=W, H=wl["rows"]) + " -- reference CPU implementation on a bounded band",
                   "name": args.workload, "band_rows": rows, "ic": "gaussian_bump"},
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": sample},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------- our arm --
STAGES = {"euler": 1, "rk2": 2, "rk4": 4}


def n_stages(wl):
    return 2 if wl["integ"] == "rk4" and wl["model"] != "shallow_water" else STAGES[wl["integ"]]


def fp32_pipe_floor(W, nrows, mhz, arith):
    """What bounds the RK4 fp32 whole-step kernel: the fp32 pipe. The per-iteration pipe
    cycles come from the committed dynamic instruction mix of the shipped kernel (profiles/instruction_mix.json,
    written by profiles/make_mix_json.py from an ncu --import-source capture; packed instructions hold the pipe for
    two cycles), strips x chunks x iterations over 148 SMs x 4 schedulers at the sampled SM clock."""
    path = os.path.join(ROOT, "profiles", "instruction_mix.json")
    try:
        with open(path) as f:
            mix = json.load(f)[arith]
    except Exception:
        return None
    strips, chunks = -(-W // mix["columns_per_strip"]), -(-nrows // mix["rows_per_chunk"])
    iters = mix["rows_per_chunk"] + 2 * 4
    cyc = mix["fma_pipe_cycles_per_iteration"]
    floor_ms = strips * chunks * iters * cyc / (148 * 4 * (mhz or 1965.0) * 1e3)
    return {"floor_ms_per_step": floor_ms, "fma_pipe_cycles_per_strip_row_iteration": cyc,
            "source": "profiles/instruction_mix.json", "sources_match_capture": mix.get("stamp") == source_stamp(),
            "note": "no-FMA fp32 work of the bit-exact arithmetic at 128 lanes/clk/SM"}


def run_b200_arm(args, wl, rank, world, local_rank):
    from weather_sim import _capi
    from weather_sim import distributed as wd

    dist = None
    if world > 1:
        import torch
        import torch.distributed as tdist
        torch.cuda.set_device(local_rank)
        tdist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        dist = tdist

    def barrier():
        if dist is not None:
            dist.barrier()

    def all_ranks(x):
        if dist is None:
            return [x]
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    def max_over_ranks(x):
        if dist is None:
            return x
        import torch
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    peak, peak_kind = measured_peaks()

    def make_sim(w, arith="strict", kernel=None):
        H_global = w["rows"] * world if w["scaling"] == "weak" else w["rows"]
        np_dtype = np.float64 if w["dtype"] == "f64" else np.float32
        nccl_id = wd.share_nccl_id() if world > 1 else None
        sim = _capi.Simulation(w["W"], H_global, model=w["model"], integrator=w["integ"], coriolis_f=0.1, max_time=1e30,
                               dtype=np_dtype, num_levels=w["levels"], kernel=kernel or args.kernel,
                               device_id=local_rank, rank=rank, nranks=world, nccl_id=nccl_id, arith=arith,
                               dx=w.get("dx", 1.0), dy=w.get("dy", 1.0), extended=w.get("extended"),
                               rk4_classical=w.get("classical", False))
        return sim, H_global, np_dtype

    def time_device(sim, steps, warmup):
        """W warm-up steps, then exactly K steps under CUDA events on the library's stepping stream (all ranks are
        aligned on the device right before the first timed step), max over ranks."""
        sim.step(warmup)
        launches0 = sim.metrics.kernel_launches
        sampler = ClockSampler(local_rank)  # NVML initialisation (milliseconds, differs per rank) stays out of the region
        barrier()
        with sampler as clk:
            sim.advance_async(steps)
            sim.synchronize()
        dev_ms = sim.last_run_device_ms
        barrier()
        return {"dev_ms": dev_ms, "ms": max_over_ranks(dev_ms), "ms_ranks": all_ranks(dev_ms),
                "launches": sim.metrics.kernel_launches - launches0, "clocks": clk.summary()}

    def roofline_of(w, sim, cells_rank, steps, dev_ms, launches):
        kernel = sim.kernel_name
        fused = kernel.startswith("step_fused")
        bpc = w["bpc_step"] if fused else w["bpc_stage"]
        achieved = bpc * cells_rank * steps / (dev_ms * 1e-3) / 1e9
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "peak_kind": peak_kind, "kernel": kernel, "algorithmic_bytes_per_cell_step": bpc,
                "launches_per_step": launches / max(steps, 1),
                "stage_per_pass_equivalent_gbs": w["bpc_stage"] * cells_rank * steps / (dev_ms * 1e-3) / 1e9}

    def sub_line(name, arith="strict", steps=None, warmup=3):
        """One more BASELINE configuration, device-resident, in the same process and on the same box."""
        w = WORKLOADS[name]
        steps = steps or max(5, min(args.steps, 20))
        sim, H_global, np_dtype = make_sim(w, arith=arith, kernel="auto")
        r0, nrows = sim.local_rows
        L = w["levels"]
        h = np.empty((nrows, w["W"]) if L == 1 else (L, nrows, w["W"]), np_dtype)
        fill_bump_fast(h[0] if L > 1 else h, r0, w["W"], H_global)
        for k in range(1, L):  # per-level amplitude (1 + k/64), SURVEY.md section 8d C4
            h[k] = 10.0 + (1.0 + k / 64.0) * (h[0] - 10.0)
        sim.set_state(h=h)  # u = v = 0 are the reset() defaults already on the device
        del h
        t = time_device(sim, steps, warmup)
        cells_total, cells_rank = w["W"] * H_global * L, w["W"] * nrows * L
        rl = roofline_of(w, sim, cells_rank, steps, t["dev_ms"], t["launches"])
        out = {"name": name + ("" if arith == "strict" else "+" + arith), "metric": w["metric"],
               "workload": w["desc"].format(W=w["W"], H=H_global), "scaling": w["scaling"],
               "value": cells_total * steps / (t["ms"] * 1e-3), "unit": UNIT, "ms_per_step": t["ms"] / steps,
               "steps": steps, "warmup": warmup, "dtype": w["dtype"], "n_gpus": world, "kernel": sim.kernel_name,
               "arith": arith, "rows_per_gpu": nrows, "gpu_launches": int(t["launches"]), "clocks": t["clocks"],
               "roofline": {k: rl[k] for k in ("bound", "achieved", "peak", "unit", "frac", "algorithmic_bytes_per_cell_step",
                                               "stage_per_pass_equivalent_gbs")}}
        if world > 1:
            out["ms_per_step_per_rank"] = [m / steps for m in t["ms_ranks"]]
        sim.close()
        return out

    # ---- the headline workload -------------------------------------------------------------------------------
    W, L = wl["W"], wl["levels"]
    sim, H_global, np_dtype = make_sim(wl, arith=args.arith)
    esize = np.dtype(np_dtype).itemsize
    r0, nrows = sim.local_rows
    shape = (nrows, W) if L == 1 else (L, nrows, W)
    hb = _capi.pinned_empty(shape, np_dtype)
    ub = _capi.pinned_empty(shape, np_dtype)
    vb = _capi.pinned_empty(shape, np_dtype)
    fill_bump_slab(hb, r0, W, H_global)
    if L > 1:  # per-level amplitude (1 + k/64), SURVEY.md section 8d C4
        for k in range(1, L):
            hb[k] = 10.0 + (1.0 + k / 64.0) * (hb[0] - 10.0)
    ub[...] = 0.0
    vb[...] = 0.0
    sim.set_state(ub, vb, hb)

    t = time_device(sim, args.steps, args.warmup)
    dev_ms, ms, launches, clocks = t["dev_ms"], t["ms"], t["launches"], t["clocks"]
    cells_total = W * H_global * L
    cells_rank = W * nrows * L
    value = cells_total * args.steps / (ms * 1e-3)
    halo_us, halo_bytes = sim.time_halo_exchange(100) if world > 1 else (0.0, 0)
    halo_us = max_over_ranks(halo_us)

    # ---- end to end through the C-ABI with HOST buffers, every step: host state in, one step, host state out.
    # Simulation.step_host streams the state through the GPU in row slabs (upload / step / download overlap); the
    # plain three-call sequence (set, step, get) is timed beside it.
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, 10))
        outs = [_capi.pinned_empty(shape, np_dtype) for _ in range(3)]

        def seq_step():
            sim.set_state(ub, vb, hb)
            sim.step(1)
            for n, o in zip(("u", "v", "h"), outs):
                sim.get_field(n, out=o)

        def host_step():
            sim.step_host(ub, vb, hb, *outs)

        def timed(fn):
            for _ in range(2):
                fn()
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                fn()
            dt = max_over_ranks(time.perf_counter() - t0)
            barrier()
            return dt

        seq_s = timed(seq_step)
        best_s = timed(host_step)
        e2e = {"value": cells_total * e2e_steps / best_s, "unit": UNIT,
               "h2d_bytes_per_step": 3 * cells_total * esize, "d2h_bytes_per_step": 3 * cells_total * esize,
               "steps": e2e_steps, "ms_per_step": best_s / e2e_steps * 1e3,
               "api": ("weather_sim._capi.Simulation.step_host (C-ABI wsb_sim_step_host): pinned host u,v,h in, one "
                       "step, pinned host u,v,h out, every step; slabs stream H2D / kernel / D2H concurrently"),
               "unpipelined_three_call_value": cells_total * e2e_steps / seq_s}
        for o in outs:
            _capi.pinned_free(o)

    roofline = clk_mhz = None
    if rank == 0:
        kernel = sim.kernel_name
        fused = kernel.startswith("step_fused")
        roofline = roofline_of(wl, sim, cells_rank, args.steps, dev_ms, launches)
        roofline["note"] = (("whole-step kernel: every RK stage stays on chip, HBM sees one read of y_n and one write of "
                             "y_n+1 per step (%d B/cell-step instead of %d with one pass per stage); for RK4 fp32 the "
                             "kernel is bound by the fp32 pipe, not HBM (no contraction allowed for bit parity) -- "
                             "DESIGN.md section 4.4" % (wl["bpc_step"], wl["bpc_stage"])) if fused else
                            "one fused tendency+update pass per RK stage")
        if fused and wl["dtype"] == "f32" and wl["integ"] == "rk4" and wl["model"] == "shallow_water":
            fl = fp32_pipe_floor(W, nrows, clocks["sm_mhz"], args.arith)
            if fl:
                fl["frac"] = fl["floor_ms_per_step"] / (dev_ms / args.steps)
                roofline["fp32_pipe"] = fl
        traffic_file = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(traffic_file):
            try:
                with open(traffic_file) as f:
                    tj = json.load(f)
                roofline["traffic"] = tj.get(args.workload, {}).get(kernel)
                roofline["traffic_sources_match_capture"] = tj.get("_stamp") == source_stamp()
            except Exception:
                pass
    kernel_name = sim.kernel_name
    sim.close()
    for b in (ub, vb, hb):
        _capi.pinned_free(b)
    del ub, vb, hb

    # ---- the call a user of the reference makes: pyweather_sim.WeatherSimulation.run(n) + one field read-back ----
    pyb = None
    if world == 1 and not args.no_e2e:
        try:
            pyb = pybind_run_leg(wl, args, cells_total)
        except Exception as exc:  # reported, never hidden
            pyb = {"error": repr(exc)}

    # ---- the other BASELINE configurations, same process, same box --------------------------------------------
    others, strong = [], None
    if not args.no_other_configs and args.workload == "swe8192_rk4":
        if world == 1:
            for name, arith in (("swe8192_euler", "strict"), ("baro16384_f64", "strict"), ("prim2048x64", "strict"),
                                ("swe8192_rk4_div", "strict"), ("swe8192_rk4", "folded"), ("swe8192_rk4_ext", "strict"),
                                ("prim2048x64_ext", "strict")):
                try:
                    others.append(sub_line(name, arith))
                except Exception as exc:  # a failing sub-line is reported, it does not take the headline with it
                    others.append({"name": name, "error": repr(exc)})
        try:
            strong = sub_line("swe32768_rk4", "strict", steps=max(5, min(args.steps, 10)))
        except Exception as exc:
            if world > 1:
                raise  # collective: every rank would have to fail alike
            strong = {"name": "swe32768_rk4", "error": repr(exc)}

    if rank == 0:
        line = {
            "metric": wl["metric"], "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": wl["scaling"], "vs_baseline": None, "dtype": wl["dtype"], "data": "synthetic",
            "config": {"workload": wl["desc"].format(W=W, H=H_global) + ", dt=0.01, g=9.81, f=0.1, Gaussian bump",
                       "name": args.workload, "grid": [H_global, W] if L == 1 else [L, H_global, W],
                       "kernel": kernel_name, "arith": args.arith,
                       "decomposition": f"row-slabs x{world} ({nrows} rows per GPU)",
                       "cache": f"state per GPU ({3 * cells_rank * esize / 1e6:.0f} MB read per step) >> 126 MB L2: "
                                "inputs larger than L2"},
            "clocks": clocks,
            "gpu_launches": int(launches),
            "roofline": roofline,
        }
        if e2e is not None:
            if pyb is not None:
                e2e["pyweather_sim_run"] = pyb
            line["e2e"] = e2e
        if others:
            line["other_configs"] = others
        if strong is not None:
            line["strong_scaling"] = strong
        if world > 1:
            line["ms_per_step_per_rank"] = [m / args.steps for m in t["ms_ranks"]]
            nv = 900.0e9  # NVLink 5 per direction per GPU
            line["halo"] = {"bytes_per_neighbour_per_step": halo_bytes,
                            "exchange_us": halo_us, "exchanges_timed": 100,
                            "achieved_gbs_per_direction": (halo_bytes / (halo_us * 1e-6) / 1e9) if halo_us else None,
                            "nvlink_bound_us": halo_bytes / nv * 1e6,
                            "frac_of_nvlink_bound": (halo_bytes / nv * 1e6 / halo_us) if halo_us else None,
                            "note": "100 bare ghost-row exchanges (the step's own ncclSend/ncclRecv group, no compute) "
                                    "timed with CUDA events on the comm stream after a device-side rendezvous, max over "
                                    "ranks; bound = bytes / 900 GB/s. At these sizes the exchange is latency-bound; "
                                    "in a step it runs on its own stream beside the interior sweep"}
        if world == 1 and not args.no_cpu_baseline:
            n = min(wl["rows"], 2048 if wl["integ"] != "euler" else 4096)
            # ~10 s of host work at the reference's 25-30 Mcell/s (RK) / 100+ Mcell/s (Euler) with a full OpenMP team
            steps = (70 if wl["integ"] != "euler" else 60) if wl["rows"] >= 2048 else 3
            code = (MODEL_CODE[wl["model"]], INTEG_CODE[wl["integ"]])
            r = time_reference_cpu(n, n, steps, 2, *code)
            r1 = time_reference_cpu(n, n, max(steps // 14, 2), 1, *code, threads=1)
            line["cpu_baseline"] = {
                "value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                "sample": f"reference CPU path, {wl['model']} {n}x{n} fp32 {wl['integ']}, {steps} steps after 2 "
                          f"warm-up ({r['seconds']:.1f} s), OpenMP team of {r['cores']} threads (measured); its "
                          "throughput is flat in grid size (BASELINE.md section 2)",
                "cpu_model": cpu_model(), "host_threads_available": host_threads(),
                "one_thread_value": r1["value"], "one_thread_seconds": r1["seconds"]}
            if args.workload == "swe8192_rk4":
                try:
                    line["cpu_baseline"]["config1"] = config1_reference_cpu()
                except Exception as e:  # a reported extra, never a reason to lose the line
                    line["cpu_baseline"]["config1"] = {"error": f"{type(e).__name__}: {e}"}
        emit(line)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def pybind_run_leg(wl, args, cells_total):
    """End to end the way a user of the reference drives it (weather_simulation.py:306-322 -> python_bindings.cpp:343,
    255): `WeatherSimulation.run(n)` on the drop-in pyweather_sim module, then ONE field read-back into numpy."""
    import weather_sim.pyweather_sim as m
    os.environ.setdefault("WSB_QUIET", "1")
    c = m.SimulationConfig()
    c.grid_width, c.grid_height = wl["W"], wl["rows"]
    c.model = {"shallow_water": m.SimulationModel.ShallowWater, "barotropic": m.SimulationModel.Barotropic,
               "primitive": m.SimulationModel.PrimitiveEquations}[wl["model"]]
    c.integration_method = {"euler": m.IntegrationMethod.ExplicitEuler, "rk2": m.IntegrationMethod.RungeKutta2,
                            "rk4": m.IntegrationMethod.RungeKutta4}[wl["integ"]]
    c.coriolis_f = 0.1
    c.max_time = 1.0e30
    c.double_precision = wl["dtype"] == "f64"
    c.num_levels = wl["levels"]
    sim = m.WeatherSimulation(c)
    sim.initialize()
    h = np.empty((wl["rows"], wl["W"]), np.float32)
    fill_bump_fast(h, 0, wl["W"], wl["rows"])
    if wl["levels"] > 1:
        h = np.broadcast_to(h, (wl["levels"],) + h.shape).copy()
    g = sim.get_current_grid()
    g.set_height_field(h)
    sim.run(args.warmup)
    g.get_height_field()
    n = args.steps
    t0 = time.perf_counter()
    sim.run(n)
    out = sim.get_current_grid().get_height_field()
    dt = time.perf_counter() - t0
    return {"value": cells_total * n / dt, "unit": UNIT, "steps": n, "ms_per_call": dt * 1e3,
            "h2d_bytes_per_call": 0, "d2h_bytes_per_call": int(out.nbytes),
            "api": "weather_sim.pyweather_sim.WeatherSimulation.run(n) + get_current_grid().get_height_field() "
                   "(pageable numpy result): the reference user's loop, wall clock around both calls"}


_JSON_OUT = None


def emit(line):
    """The ONE JSON line of the contract, on the real stdout."""
    out = _JSON_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    # stdout carries exactly one JSON line: everything else that writes to fd 1 (the NCCL version banner, the
    # reference module's "Completed N steps" chatter) is sent to stderr
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="swe8192_rk4", choices=sorted(WORKLOADS))
    ap.add_argument("--kernel", default="auto", choices=["auto", "stage_direct", "step_fused_reg", "step_fused_tma"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="tuning runs only: skip the host-buffer end-to-end leg")
    ap.add_argument("--no-other-configs", action="store_true",
                    help="skip the sub-lines for the other BASELINE configurations (other_configs, strong_scaling)")
    ap.add_argument("--arith", default="strict", choices=["strict", "folded"],
                    help="fp32 evaluation of the packed whole-step kernels (include/weather_b200.h wsb_arith_mode)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference_arm(args, wl, rank)
        return
    if world == 1 and args.gpus > 1:
        print(f"bench.py: --gpus {args.gpus} needs torchrun (WORLD_SIZE=1 here)", file=sys.stderr)
        sys.exit(2)
    run_b200_arm(args, wl, rank, world, local_rank)


if __name__ == "__main__":
    main()
