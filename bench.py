#!/usr/bin/env python
"""bench.py -- scattering-moment evaluations per second on B200 (BASELINE.json metric).

A "step" is one pass of the hot path over one nuclide: calc_elastic_grid + calc_inelastic_grid of the
U-238-shape synthetic nuclide (BASELINE.json configs[1]: elastic + 40 discrete levels + Law-44
continuum, P0-P7, 70 groups, 20 000 E_in points).  One evaluation = one output element (E_in, g, l).

  value     device-resident: tables, E_in and the moment arrays stay in HBM; CUDA-event time on the
            library's stream, max over ranks.
  e2e       the same work through the reference-facing call calc_scatt(...) with HOST buffers:
            table upload + convert_distro + integration + D2H of the matrices, every step.
  roofline  dominant kernel k_file6_cm_ws (integrate_file6_cm_leg): algorithmic FP64 flops (SURVEY 8d
            formula F_B) / its CUDA-event time, against the FP64 FMA rate measured in this
            run by ndppgpu_measure_fp64_peak (MEASURED_PEAKS.json has no FP64 figure).  The path is
            FP64-pipe bound, so `bound` is "fp64" (neither of the contract's hbm | tensor); the HBM view is
            reported beside it.
  cpu_baseline / --impl reference
            the CPU oracle (restatement of the reference algorithm; the image has no Fortran compiler)
            on all host cores, on a bounded sample of the same E_in grids.

N > 1 (torchrun, one rank per GPU): weak scaling -- every rank integrates its own nuclide of the same
shape (different seed), i.e. the library is sharded by nuclide; the moment arrays are gathered to
rank 0 over NCCL inside the timed region, as the reference's driver needs them before output.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "scatt moment evals/s (E_in x group x l)"
UNIT = "evals/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-grid", type=int, default=20000, help="E_in points of the C2 nuclide (default: the named 20k)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-sample", type=int, default=0, help="E_in points in the CPU sample (0 = auto)")
    return ap.parse_args()


# ---- workload ------------------------------------------------------------------------------------
def make_workload(n_grid, seed_offset=0):
    from ndpp_b200 import synth
    nuc, e_bins, params, Ein_el, Ein_inel = synth.c2_u238(n_grid=n_grid, seed=synth.SEED0 + 2 + 1000 * seed_offset)
    return nuc, e_bins, params, Ein_el, Ein_inel


def workload_config(nuc, e_bins, params, Ein_el, Ein_inel, n_gpus):
    return {"workload": "C2 U-238-shape synthetic ACE nuclide: elastic + 40 levels (law 3) + continuum (law 44, CM)",
            "groups": len(e_bins) - 1, "legendre_orders": params.order + 1, "mu_bins": params.mu_bins,
            "ne_per_grp": params.ne_per_grp, "NE_elastic": int(len(Ein_el)), "NE_inelastic": int(len(Ein_inel)),
            "reactions": len(nuc.reactions), "nuclides": n_gpus,
            "parallelism": f"nuclide-sharded x{n_gpus}" if n_gpus > 1 else "single GPU",
            "l2": "flushed between steps (256 MiB write)"}


def evals_per_step(e_bins, params, Ein_el, Ein_inel):
    GL = (len(e_bins) - 1) * (params.order + 1)
    return (len(Ein_el) + len(Ein_inel)) * GL * 1  # nuscatter is off for C2


def file6_cm_flops(nuc, e_bins, params, Ein_inel):
    """Algorithmic flops of integrate_file6_cm_leg + unitbase for the continuum reaction over the
    inelastic grid (F_B): per active E_in
        11*M*NPu + G_b*K*[17 + 71*M + (15L+11)(M-1)] + 3*G_b*L
    with G_b = g_hi - g_lo + 1 evaluated from the reference's own bounds (scattdata_header.F90:1138-1166)."""
    from ndpp_b200.ace import N_NC
    M, L, K = params.mu_bins, params.order + 1, params.ne_per_grp
    rx = [r for r in nuc.reactions if r.MT == N_NC]
    if not rx:
        return 0.0, 0
    r = rx[0]
    d = r.edist.data
    NE = int(d[1])
    e_in = d[2:2 + NE]
    locs = d[2 + NE:2 + 2 * NE].astype(int)
    NP = np.array([int(d[lc + 1]) for lc in locs])
    last = np.array([d[lc + 2 + NP[i] - 1] for i, lc in enumerate(locs)])
    E = Ein_inel[(Ein_inel > nuc.energy[r.threshold - 1]) & (Ein_inel <= e_bins[-1])]
    iE = np.clip(np.searchsorted(e_in, E, side="right") - 1, 0, NE - 2)
    f = (E - e_in[iE]) / (e_in[iE + 1] - e_in[iE])
    eout_last = (1 - f) * last[iE] + f * last[iE + 1]
    awr = nuc.awr
    Eo_hi = eout_last + (E + 2 * (awr + 1) * np.sqrt(E * eout_last)) / (awr + 1) ** 2
    g_lo = np.searchsorted(e_bins, 1e-12, side="right") - 1
    g_hi = np.where(Eo_hi >= e_bins[-1], len(e_bins) - 2, np.searchsorted(e_bins, Eo_hi, side="right") - 1)
    Gb = g_hi - g_lo + 1
    NPu = NP[iE] + NP[iE + 1] - 1
    flops = 11.0 * M * NPu + Gb * K * (17.0 + 71.0 * M + (15.0 * L + 11.0) * (M - 1)) + 3.0 * Gb * L
    return float(flops.sum()), int(len(E))


# ---- clocks --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.samples, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            try:
                sm.append(float(p[0])); mx.append(float(p[1]))
            except Exception:
                continue
            for n, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ---- CPU arm (oracle) ------------------------------------------------------------------------------
def cpu_run(nuc, e_bins, params, Ein_el, Ein_inel, n_sample, threads, seed=0):
    """Times the oracle on a shuffled, evenly spread sample of both E_in grids.  Returns
    (evals/s, seconds, description)."""
    from oracle import pyoracle
    rng = np.random.default_rng(seed)
    frac = min(1.0, n_sample / float(len(Ein_el) + len(Ein_inel)))
    s_el = rng.permutation(Ein_el[::max(1, int(round(1 / frac)))])
    s_in = rng.permutation(Ein_inel[::max(1, int(round(1 / frac)))])
    pyoracle.lib().ref_set_omp_chunk(1)  # the sample is small: chunks of 100 would serialise it
    t0 = time.perf_counter()
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    t_conv = time.perf_counter() - t0
    t0 = time.perf_counter()
    rn.elastic(s_el, n_threads=threads)
    rn.inelastic(s_in, n_threads=threads)
    dt = time.perf_counter() - t0
    rn.close()
    GL = (len(e_bins) - 1) * (params.order + 1)
    ev = (len(s_el) + len(s_in)) * GL
    # convert_distro is once per nuclide: charge it in proportion to the sampled fraction
    total = dt + t_conv * frac
    desc = (f"{len(s_el)} of {len(Ein_el)} elastic + {len(s_in)} of {len(Ein_inel)} inelastic E_in (every "
            f"{max(1, int(round(1 / frac)))}th point, shuffled, dynamic chunk 1), {threads} OpenMP threads, "
            f"oracle = C restatement of the reference (gcc -O3, no FMA contraction)")
    return ev / total, total, desc


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nuc, e_bins, params, Ein_el, Ein_inel = make_workload(args.n_grid)
    n_sample = args.cpu_sample or 1200
    for _ in range(args.warmup):
        cpu_run(nuc, e_bins, params, Ein_el, Ein_inel, max(64, n_sample // 16), threads)
    vals, secs, desc = [], [], ""
    for k in range(args.steps):
        v, s, desc = cpu_run(nuc, e_bins, params, Ein_el, Ein_inel, n_sample, threads, seed=k)
        vals.append(v); secs.append(s)
    value = float(np.sum([v * s for v, s in zip(vals, secs)]) / np.sum(secs))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(secs)),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(nuc, e_bins, params, Ein_el, Ein_inel, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ---- GPU arm -------------------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist
    from ndpp_b200 import scatt
    from ndpp_b200.capi import Context

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nuc, e_bins, params, Ein_el, Ein_inel = make_workload(args.n_grid, seed_offset=rank)
    ctx = Context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    GL = (len(e_bins) - 1) * (params.order + 1)
    ev_step = evals_per_step(e_bins, params, Ein_el, Ein_inel)

    # device-resident state
    dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
    d_Eel = torch.from_numpy(Ein_el).to(dev)
    d_Ein = torch.from_numpy(Ein_inel).to(dev)
    n_el, n_in = len(Ein_el), len(Ein_inel)
    max_el, max_in = n_el, n_in
    if world > 1:
        # the ranks' nuclides have the same shape but np.unique may drop a duplicate grid point on some
        # of them: pad every slab to the largest so that one NCCL gather per matrix assembles them
        sizes = torch.tensor([n_el, n_in], device=dev)
        dist.all_reduce(sizes, op=dist.ReduceOp.MAX)
        max_el, max_in = int(sizes[0]), int(sizes[1])
    d_el_pad = torch.zeros((max_el, GL), dtype=torch.float64, device=dev)
    d_inel_pad = torch.zeros((max_in, GL), dtype=torch.float64, device=dev)
    d_el, d_inel = d_el_pad[:n_el], d_inel_pad[:n_in]
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)
    gather_el = gather_in = None
    if world > 1 and rank == 0:
        gather_el = [torch.empty_like(d_el_pad) for _ in range(world)]
        gather_in = [torch.empty_like(d_inel_pad) for _ in range(world)]

    def step_device():
        with torch.cuda.stream(lib_stream):
            flush.fill_(0.0)
        dn.elastic_dev(d_Eel, d_el)
        dn.inelastic_dev(d_Ein, d_inel)
        if world > 1:
            torch.cuda.current_stream().wait_stream(lib_stream)
            dist.gather(d_el_pad, gather_el, dst=0)
            dist.gather(d_inel_pad, gather_in, dst=0)
            lib_stream.wait_stream(torch.cuda.current_stream())

    def timed(fn, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(steps):
            fn()
        torch.cuda.current_stream().wait_stream(lib_stream)
        lib_stream.wait_stream(torch.cuda.current_stream())
        e1.record(lib_stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = time.perf_counter() - t0
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall * 1e3], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0]), float(t[1]) / 1e3
        return ms, wall

    for _ in range(args.warmup):
        step_device()
    torch.cuda.synchronize()
    ctx.stats(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall = timed(step_device, args.steps)
    st = ctx.stats(reset=True)
    clk = clocks.stop() if rank == 0 else None
    # the flush kernel is inside the event bracket: subtract its measured cost
    with torch.cuda.stream(lib_stream):
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(lib_stream)
        for _ in range(args.steps):
            flush.fill_(0.0)
        f1.record(lib_stream)
    torch.cuda.synchronize()
    flush_ms = f0.elapsed_time(f1)
    ms_step = (ms - flush_ms) / args.steps
    value = world * ev_step / (ms_step * 1e-3)

    # ---- e2e: calc_scatt with host (pinned) buffers ----------------------------------------------
    # page-locked host buffers for the E_in grids and the result matrices (the contract's e2e path)
    h_Eel = torch.from_numpy(Ein_el).pin_memory().numpy()
    h_Ein = torch.from_numpy(Ein_inel).pin_memory().numpy()
    h_el = torch.empty((len(Ein_el), GL), dtype=torch.float64).pin_memory().numpy()
    h_inel = torch.empty((len(Ein_inel), GL), dtype=torch.float64).pin_memory().numpy()

    def step_e2e():
        dn2 = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        el = dn2.elastic(h_Eel, out=h_el)
        inel, _ = dn2.inelastic(h_Ein, False, out=h_inel)
        dn2.clear()
        return el, inel

    step_e2e()
    ctx.stats(reset=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    st2 = ctx.stats(reset=True)
    e2e_value = world * ev_step * args.steps / e2e_s

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    flops, n_act = file6_cm_flops(nuc, e_bins, params, Ein_inel)
    f6_ms = st["file6_cm_ms"] / max(1, st["file6_cm_launches"])
    peak = ctx.measure_fp64_peak(0.5)
    achieved = flops / (f6_ms * 1e-3) / 1e12 if f6_ms > 0 else 0.0
    traffic, ncu_extra = None, None
    try:   # DRAM bytes of one launch of the dominant kernel from the committed ncu --set full capture
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        if args.n_grid == 20000:
            traffic = tj["traffic_bytes_per_launch"]
            ncu_extra = {k: tj[k] for k in ("fp64_pipe_active_pct", "achieved_occupancy_pct", "issue_slots_busy_pct",
                                            "duration_ms_under_ncu", "fp64_issue_floor_ms", "source") if k in tj}
    except Exception:
        pass
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
            "kernel": "k_file6_cm_ws (integrate_file6_cm_leg: records + tables + pipeline kernel)",
            "kernel_ms": f6_ms, "kernel_share_of_step": f6_ms / ms_step if ms_step else None,
            "algorithmic_flops_per_launch": flops, "active_E_in": n_act,
            "peak_source": "measured in this run: ndppgpu_measure_fp64_peak (DFMA chains, all SMs); "
                           "MEASURED_PEAKS.json has no FP64 figure"}
    if ncu_extra:
        roof["ncu"] = ncu_extra    # counters of the committed ncu --set full capture of this kernel on this workload
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        out_bytes = 8.0 * ev_step
        roof["hbm"] = {"algorithmic_bytes_per_step": out_bytes, "achieved_gbs": out_bytes / (ms_step * 1e-3) / 1e9,
                       "peak_gbs": peaks.get("hbm_gbs"), "note": "output-write bytes only; the path is FP64-bound"}
    except Exception:
        pass

    cpu = None
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        v, s, desc = cpu_run(nuc, e_bins, params, Ein_el, Ein_inel, args.cpu_sample or 1200, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc, "seconds": s}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": workload_config(nuc, e_bins, params, Ein_el, Ein_inel, world),
                "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st2["h2d_bytes"] / args.steps,
                        "d2h_bytes_per_step": st2["d2h_bytes"] / args.steps,
                        "call": "ndpp_b200.scatt.DeviceNuclide(...) + elastic(host) + inelastic(host) == calc_scatt"},
                "gpu_launches": int(st["launches"]), "roofline": roof, "cpu_baseline": cpu,
                "wall_s_timed_region": wall, "flush_ms_subtracted": flush_ms}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
