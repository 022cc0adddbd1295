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

N > 1 (torchrun, one rank per GPU): STRONG scaling of the same nuclide -- its E_in grids are dealt cyclically
over the GPUs inside libndppgpu.so (ndppgpu_group_*, csrc/group.cuh), every rank integrates its columns, and the
columns are gathered to the root device with ncclSend / ncclRecv on a side stream (two result buffers in turn,
so the gather overlaps the next step's kernels).  torch.distributed only carries the NCCL id, the barrier and the
max over ranks.  `e2e` at N > 1 includes the table uploads on every GPU, the gather and the D2H on the root.

extra.configs     C1, C3 (293.6 / 600 / 1200 K), C4 (discrete, continuous, 16-bin histogram): evals/s, kernel ms,
                  roofline fraction from the algorithmic work of SURVEY 8d, sampled CPU rate (N = 1)
extra.c5_library  the 300-nuclide library (BASELINE configs[4]) through ndppgpu_plan_library / ndppgpu_library_run
                  on the same N GPUs: seconds, evals/s, measured and modelled imbalance
extra.strong_scaling_breakdown   per-rank kernel time against the step (N > 1)
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
    ap.add_argument("--no-extras", action="store_true", help="skip extra.configs (C1, C3, C4) and extra.c5_library")
    ap.add_argument("--c5-nuclides", type=int, default=300, help="nuclides of the C5 library in extra.c5_library")
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
            "reactions": len(nuc.reactions), "nuclides": 1,
            "parallelism": (f"one nuclide, E_in dealt cyclically over {n_gpus} GPUs inside libndppgpu.so (ndppgpu_group_*), "
                            "NCCL gather to the root on a side stream") if n_gpus > 1 else "single GPU",
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
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(nuc, e_bins, params, Ein_el, Ein_inel, args.gpus),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            # a step of this arm is a SAMPLE of the configuration (every k-th E_in), not the configuration: `value` is the
            # rate on the sample, ms_per_step the sample's time; the whole configuration at this rate would take:
            "sampled": True,
            "ms_per_step_full_config_extrapolated": 1e3 * evals_per_step(e_bins, params, Ein_el, Ein_inel) / value}
    print(json.dumps(line))


# ---- GPU arm -------------------------------------------------------------------------------------
def _flush_cost(torch, lib_stream, flush, steps):
    with torch.cuda.stream(lib_stream):
        f0 = torch.cuda.Event(enable_timing=True); f1 = torch.cuda.Event(enable_timing=True)
        f0.record(lib_stream)
        for _ in range(steps):
            flush.fill_(0.0)
        f1.record(lib_stream)
    torch.cuda.synchronize()
    return f0.elapsed_time(f1)


def run_b200(args):
    import torch
    import torch.distributed as dist
    from ndpp_b200 import library, scatt
    from ndpp_b200.capi import Context
    from ndpp_b200.group import Group, GroupNuclide

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        # torch.distributed carries only what an MPI driver would carry over MPI: the NCCL id, the barrier and the
        # max over ranks of the timings.  Sharding, NCCL communicator and gather are the library's (csrc/group.cuh).
        dist.init_process_group("nccl", device_id=dev)
        group = Group.from_rank(local, rank, world, library.broadcast_id(dist, dev))
        ctx = group.ctx(0)
    else:
        ctx = Context(local)

    nuc, e_bins, params, Ein_el, Ein_inel = make_workload(args.n_grid)      # the same nuclide on every rank
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    GL = (len(e_bins) - 1) * (params.order + 1)
    ev_step = evals_per_step(e_bins, params, Ein_el, Ein_inel)
    n_el, n_in = len(Ein_el), len(Ein_inel)
    flush = torch.empty(256 * 1024 * 1024 // 8, dtype=torch.float64, device=dev)

    if world == 1:
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        d_Eel = torch.from_numpy(Ein_el).to(dev)
        d_Ein = torch.from_numpy(Ein_inel).to(dev)
        d_el = torch.zeros((n_el, GL), dtype=torch.float64, device=dev)
        d_inel = torch.zeros((n_in, GL), dtype=torch.float64, device=dev)

        def step_device():
            with torch.cuda.stream(lib_stream):
                flush.fill_(0.0)
            dn.elastic_dev(d_Eel, d_el)
            dn.inelastic_dev(d_Ein, d_inel)

        def join():
            pass
    else:
        # strong scaling: the E_in grids of the one nuclide dealt cyclically over the GPUs; every step ends with the
        # NCCL gather of the columns to the root device, issued on a side stream with two result buffers in turn, so
        # that it overlaps the kernels of the next step
        gn = GroupNuclide(nuc, e_bins, params, group)
        gn.set_grids(Ein_el, Ein_inel)

        def step_device():
            with torch.cuda.stream(lib_stream):
                flush.fill_(0.0)
            gn.integrate(3)

        def join():
            gn.join()

    def timed(fn, steps):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(lib_stream)
        for _ in range(steps):
            fn()
        join()                      # the last gather and assembly lie inside the event bracket
        e1.record(lib_stream)
        torch.cuda.synchronize()
        if world > 1:
            gn.sync()
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
    join()
    torch.cuda.synchronize()
    if world > 1:
        gn.sync()
    ctx.stats(reset=True)
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
    ms, wall = timed(step_device, args.steps)
    st = ctx.stats(reset=True)
    clk = clocks.stop() if rank == 0 else None
    flush_ms = _flush_cost(torch, lib_stream, flush, args.steps)    # the flush is inside the event bracket: subtract it
    ms_step = (ms - flush_ms) / args.steps
    value = ev_step / (ms_step * 1e-3)          # whole job: one nuclide, whatever the number of GPUs

    breakdown = None
    if world > 1:
        # per-phase picture of a step, per rank: what the device did (CUDA-event kernel time) against the step
        mine = torch.tensor([st["kernel_ms"] / args.steps, st["file6_cm_ms"] / args.steps, st["host_call_ms"] / args.steps,
                             st["host_sync_ms"] / args.steps, float(st["launches"]) / args.steps], dtype=torch.float64, device=dev)
        allr = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        allr = [a.tolist() for a in allr]
        k = [a[0] for a in allr]
        breakdown = {"ms_per_step": ms_step, "kernel_ms_per_rank": k, "file6_cm_ms_per_rank": [a[1] for a in allr],
                     "host_call_ms_per_rank": [a[2] for a in allr], "host_sync_ms_per_rank": [a[3] for a in allr],
                     "launches_per_rank": [a[4] for a in allr],
                     "kernel_imbalance": max(k) / (sum(k) / world),
                     "ms_outside_kernels_slowest_rank": ms_step - max(k),
                     "gathered_bytes_per_step": group.gathered_bytes() / (args.steps + args.warmup) if rank == 0 else None,
                     "note": "kernel_ms = CUDA-event time of the integrator kernels of a rank (they run back to back on "
                             "one stream); ms_per_step - max(kernel_ms) = launch gaps, host read-backs (n_act, error "
                             "latch) and whatever of the gather is not hidden behind the next step"}

    # ---- e2e: the reference-facing call with HOST buffers ----------------------------------------------------------
    h_Eel = torch.from_numpy(Ein_el).pin_memory().numpy()
    h_Ein = torch.from_numpy(Ein_inel).pin_memory().numpy()
    h_el = torch.empty((n_el, GL), dtype=torch.float64).pin_memory().numpy() if rank == 0 else None
    h_inel = torch.empty((n_in, GL), dtype=torch.float64).pin_memory().numpy() if rank == 0 else None

    if world == 1:
        def step_e2e():
            dn2 = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
            dn2.calc(h_Eel, h_Ein, False, el_out=h_el, inel_out=h_inel)
            dn2.clear()
        e2e_call = ("ndpp_b200.scatt.DeviceNuclide(...) + calc(host grids, host matrices) == calc_scatt (ndppgpu_calc_scatt: "
                    "the elastic matrices leave for the host while the inelastic kernels run)")
    else:
        def step_e2e():   # table upload + convert_distro on every GPU, deal of the grids, integration, NCCL gather, D2H on the root
            g2 = GroupNuclide(nuc, e_bins, params, group)
            g2.elastic(h_Eel, out=h_el)
            g2.inelastic(h_Ein, out=h_inel)
            g2.clear()
        e2e_call = ("ndpp_b200.group.GroupNuclide(...) + elastic(host) + inelastic(host) == calc_scatt on the device group "
                    "(ndppgpu_group_*: includes the NCCL gather and the D2H of the assembled matrices on the root)")

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
    if world > 1:   # bytes over PCIe, all ranks
        t = torch.tensor([st2["h2d_bytes"], st2["d2h_bytes"]], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        st2["h2d_bytes"], st2["d2h_bytes"] = float(t[0]), float(t[1])
    e2e_value = ev_step * args.steps / e2e_s

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    flops, n_act = file6_cm_flops(nuc, e_bins, params, Ein_inel)
    f6_ms = st["file6_cm_ms"] / max(1, st["file6_cm_launches"])
    peak = ctx.measure_fp64_peak(0.5)
    if world > 1:       # this rank integrated every world-th column
        flops /= world
    achieved = flops / (f6_ms * 1e-3) / 1e12 if f6_ms > 0 else 0.0
    traffic, ncu_extra = ncu_traffic(args)
    roof = {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
            "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_unit": "bytes/launch (ncu dram read+write)",
            "kernel": "k_file6_cm_ws (integrate_file6_cm_leg: records + tables + pipeline kernel)" + (" on rank 0" if world > 1 else ""),
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

    # ---- extras: the other BASELINE configurations under the driver's clock ------------------------------------------
    extra = {}
    if not args.no_extras:
        if world == 1:
            dn.clear()
            extra["configs"] = run_configs(ctx, peak, sample_cpu=not args.no_cpu_baseline)
            extra["ein_grid"] = run_ein_grid(ctx, nuc, e_bins, params, sample_cpu=not args.no_cpu_baseline)
            cgroup = Group(1, devices=[local])
            extra["c5_library"] = library.run_c5(cgroup, args.c5_nuclides)
            cgroup.close()
        else:
            gn.clear()
            extra["configs"] = run_configs_group(group, ctx, peak, dist, dev)
            extra["c5_library"] = library.run_c5(group, args.c5_nuclides, dist=dist, device=dev)
        if breakdown:
            extra["strong_scaling_breakdown"] = breakdown

    if rank == 0:
        cfg = workload_config(nuc, e_bins, params, Ein_el, Ein_inel, world)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clk,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": st2["h2d_bytes"] / args.steps,
                        "d2h_bytes_per_step": st2["d2h_bytes"] / args.steps, "call": e2e_call},
                "gpu_launches": int(st["launches"]), "roofline": roof, "cpu_baseline": cpu,
                "wall_s_timed_region": wall, "flush_ms_subtracted": flush_ms, "extra": extra}
        print(json.dumps(line))
    if world > 1:
        group.close()
        dist.destroy_process_group()


def ncu_traffic(args):
    """DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture of this workload.  The
    capture names the kernel source it was taken on (sha256 of csrc/kernels_file6_ws.cuh); when the source has changed
    since, the figure is reported as stale instead of being passed off as current."""
    import hashlib
    for name in ("r2_ncu_traffic.json", "r1_ncu_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", name)))
        except Exception:
            continue
        if args.n_grid != 20000:
            return None, None
        src = os.path.join(ROOT, "ndpp_b200", "csrc", "kernels_file6_ws.cuh")
        sha = hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
        extra = {k: tj[k] for k in ("fp64_pipe_active_pct", "achieved_occupancy_pct", "issue_slots_busy_pct",
                                    "duration_ms_under_ncu", "fp64_issue_floor_ms", "source") if k in tj}
        extra["capture"] = "profiles/" + name
        extra["kernel_source_sha16"] = sha
        extra["stale"] = tj.get("kernel_source_sha16") != sha
        return (None if extra["stale"] else tj["traffic_bytes_per_launch"]), extra
    return None, None


# ---- the other configurations (C1, C3 x 3 temperatures, C4 x 3) on one GPU -------------------------------------------------
def _cpu_rate_nuclide(nuc, e_bins, params, Ein_el, Ein_inel, n, threads, counters=False):
    """Oracle rate on n evenly spread E_in of each grid (rank 0, bounded)."""
    from oracle import pyoracle
    pyoracle.lib().ref_set_omp_chunk(1)
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    top = e_bins[-1]
    pick = lambda E: E[E <= top][:: max(1, len(E[E <= top]) // n)][:n]
    if counters:
        pyoracle.freegas_counters(reset=True)
    t0 = time.perf_counter()
    ev = rn.elastic(pick(Ein_el), n_threads=threads).size
    if Ein_inel is not None and len(Ein_inel):
        a, b = rn.inelastic(pick(Ein_inel), n_threads=threads)
        ev += a.size + (b.size if b is not None else 0)
    dt = time.perf_counter() - t0
    cnt = pyoracle.freegas_counters(reset=True) if counters else None
    rn.close()
    return ev / dt, len(pick(Ein_el)), cnt


def run_configs_group(group, ctx, peak_tflops, dist, dev):
    """The nuclide configurations other than the headline one on the device group (N > 1): C1 and C3 at three
    temperatures through ndppgpu_group_* -- the E_in grid dealt cyclically over the GPUs, the columns gathered to the
    root -- second pass of each.  evals/s from the host clock around integrate + sync (max over ranks: kernels, gather
    and assembly inside), kernel ms = the slowest rank's CUDA-event time, roofline against N times the measured peak.
    (C4, S(a,b), has no sharded form: 5e4 E_in take 2 ms on one GPU; it is in the N = 1 line.)"""
    import torch
    from ndpp_b200 import synth
    from ndpp_b200.group import GroupNuclide
    world = dist.get_world_size()

    def allred(x, op):
        t = torch.tensor([float(x)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return float(t.item())

    rows = []

    def row(name, nuc, e_bins, params, Eel, Einel, flops=None, freegas=False):
        gn = GroupNuclide(nuc, e_bins, params, group)
        gn.set_grids(Eel, Einel)
        what = 1 | (2 if Einel is not None else 0)
        for _ in range(2):
            torch.cuda.synchronize()
            dist.barrier()
            ctx.stats(reset=True)
            t0 = time.perf_counter()
            gn.integrate(what)
            gn.sync()
            wall = time.perf_counter() - t0
        st = ctx.stats(reset=True)
        gn.clear()
        wall = allred(wall, dist.ReduceOp.MAX)
        kms = allred(st["kernel_ms"], dist.ReduceOp.MAX)
        G, L = len(e_bins) - 1, params.order + 1
        ev = (len(Eel) + (len(Einel) if Einel is not None else 0)) * G * L
        r = {"config": name, "n_gpus": world, "evals": int(ev), "evals_per_s": ev / wall, "ms": wall * 1e3,
             "kernel_ms_slowest_rank": kms, "evals_per_s_kernel": ev / (kms * 1e-3)}
        if freegas:
            nk = allred(st["freegas_kernel_evals"], dist.ReduceOp.SUM)
            ns = allred(st["freegas_sab_evals"], dist.ReduceOp.SUM)
            flops = 40.0 * nk + 35.0 * ns
            r["kernel_evals"] = int(nk)
        if flops:
            ach = flops / (kms * 1e-3) / 1e12
            r["roofline"] = {"bound": "fp64", "algorithmic_flops": flops, "achieved": ach, "peak": peak_tflops * world,
                             "unit": "TFLOP/s", "frac": ach / (peak_tflops * world),
                             "formula": "as extra.configs of the N = 1 line; peak = N x the measured FP64 peak"}
        rows.append(r)

    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(997)
    G, L, M = len(e_bins) - 1, params.order + 1, params.mu_bins
    FA = 12 + 22 * G + 40 * 2 + (M + 2) * (13 + 4 * (L - 2) + 10 * L)
    FC = 2 * M * (13 + 2 * G) + G * (M - 1) * (15 * L + 11)
    row("C1 tests/test_scatt fixture (MT 51/52 relabelled), 1000 E_in", nuc, e_bins, params, Ein, Ein,
        flops=float(len(Ein) * (2 * FA + 2 * FA + FC)))
    for kT, T in ((synth.KT_293K, 293.6), (synth.KT_600K, 600), (synth.KT_1200K, 1200)):
        nuc, e_bins, params, Ein = synth.c3_h1_freegas(kT=kT)
        row(f"C3 H-1 free gas {T} K, P3, 70 groups, 1000 E_in", nuc, e_bins, params, Ein, None, freegas=True)
    return rows


def run_ein_grid(ctx, nuc, e_bins, params, sample_cpu=True):
    """Row N3: create_Ein_grid (src/scatt.F90:166-536) of the workload's nuclide on the device (second pass; the grids
    stay on the device, the timed call returns their lengths) beside the oracle's literal chain of merges on one host
    core (the routine is serial in the reference), and whether the two grids agree in every bit."""
    from ndpp_b200 import scatt
    dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
    for _ in range(2):
        ctx.stats(reset=True)
        t0 = time.perf_counter()
        (p_el, n_el), inel, status = dn.create_ein_grid(host=False)
        wall = time.perf_counter() - t0
    st = ctx.stats(reset=True)
    row = {"routine": "create_Ein_grid (src/scatt.F90:166-536)", "n_el": n_el, "n_inel": inel[1] if inel else 0,
           "device_ms": wall * 1e3, "kernel_ms": st["kernel_ms"], "launches": int(st["launches"]), "status": status}
    if sample_cpu:
        from oracle import pyoracle
        el, inl, _ = dn.create_ein_grid()
        rn = pyoracle.RefNuclide(nuc, e_bins, params)
        t0 = time.perf_counter()
        r_el, r_inl = rn.create_ein_grid()
        row["cpu"] = {"ms": (time.perf_counter() - t0) * 1e3, "cores": 1, "kind": "port"}
        rn.close()
        row["bit_identical_to_oracle"] = bool(np.array_equal(el, r_el) and (inl is None) == (r_inl is None)
                                              and (inl is None or np.array_equal(inl, r_inl)))
    dn.clear()
    return row


def run_configs(ctx, peak_tflops, sample_cpu=True):
    """evals/s, kernel time, roofline fraction (algorithmic flops of SURVEY 8d / CUDA-event kernel time / measured FP64
    peak; HBM GB/s for the memory-bound stage 2 of the continuous S(a,b) path) and the sampled CPU rate of every
    BASELINE configuration that is not the headline one.  Second pass of each (the first loads kernels and grows the pool)."""
    from ndpp_b200 import ace, scatt, synth
    threads = os.cpu_count() or 1
    rows = []
    try:
        hbm_peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs")
    except Exception:
        hbm_peak = None

    def nuclide_row(name, nuc, e_bins, params, Eel, Einel, flops, n_cpu, flop_note, counters=False):
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        for _ in range(2):
            ctx.stats(reset=True)
            t0 = time.perf_counter()
            el = dn.elastic(Eel)
            inel = dn.inelastic(Einel) if Einel is not None else None
            wall = time.perf_counter() - t0
        st = ctx.stats(reset=True)
        dn.clear()
        ev = el.size + (sum(x.size for x in inel if x is not None) if inel is not None else 0)
        row = {"config": name, "evals": int(ev), "evals_per_s_e2e": ev / wall, "kernel_ms": st["kernel_ms"],
               "evals_per_s_kernel": ev / (st["kernel_ms"] * 1e-3), "launches": int(st["launches"])}
        cnt = None
        if sample_cpu:
            rate, n, cnt = _cpu_rate_nuclide(nuc, e_bins, params, Eel, Einel, n_cpu, threads, counters)
            row["cpu"] = {"evals_per_s": rate, "cores": threads, "kind": "port", "sample": f"{n} evenly spread E_in per grid"}
        f = flops(cnt) if callable(flops) else flops
        if counters and st["freegas_kernel_evals"]:
            # the device shares kernel values between the Legendre orders of a cell: claim what it evaluated (SURVEY 8d:
            # "a kernel that legitimately shares evaluations across l may not claim the unshared count")
            row["reference_unshared_flops"] = f
            f = 40.0 * st["freegas_kernel_evals"] + 35.0 * st["freegas_sab_evals"]
            row["kernel_evals"] = int(st["freegas_kernel_evals"])
            row["sab_evals"] = int(st["freegas_sab_evals"])
            flop_note = ("F_E = 40 N_kernel + 35 N_sab with the evaluations the device performed (counted in the kernel; one "
                         "kernel value serves the orders of a group); reference_unshared_flops = the same formula with the "
                         "reference's per-order counts from the instrumented oracle on the CPU sample, scaled to the grid")
        if f:
            row["roofline"] = {"bound": "fp64", "algorithmic_flops": f, "achieved": f / (st["kernel_ms"] * 1e-3) / 1e12,
                               "peak": peak_tflops, "unit": "TFLOP/s", "frac": f / (st["kernel_ms"] * 1e-3) / 1e12 / peak_tflops,
                               "formula": flop_note}
        return row

    # C1: tests/test_scatt fixture
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(997)
    G, L, M = len(e_bins) - 1, params.order + 1, params.mu_bins
    FA = 12 + 22 * G + 40 * 2 + (M + 2) * (13 + 4 * (L - 2) + 10 * L)
    FC = 2 * M * (13 + 2 * G) + G * (M - 1) * (15 * L + 11)
    rows.append(nuclide_row("C1 tests/test_scatt fixture (MT 51/52 relabelled), 1000 E_in", nuc, e_bins, params, Ein, Ein,
                            float(len(Ein) * (2 * FA + 2 * FA + FC)), 200,
                            "per E_in: F_A (elastic) + F_A (CM level) + F_C (lab Law 44), M_act = M"))
    # C3: H-1 free gas at three temperatures
    for kT, T in ((synth.KT_293K, 293.6), (synth.KT_600K, 600), (synth.KT_1200K, 1200)):
        nuc, e_bins, params, Ein = synth.c3_h1_freegas(kT=kT)
        L = params.order + 1
        n_cpu = 12
        fl = (lambda cnt, NE=len(Ein), n=n_cpu, L=L: float(NE / n * (40.0 * cnt[0] + 35.0 * cnt[1] + 4.0 * (L - 2) * cnt[0]))
              if cnt else None)
        rows.append(nuclide_row(f"C3 H-1 free gas {T} K, P3, 70 groups, 1000 E_in", nuc, e_bins, params, Ein, None, fl, n_cpu,
                                "F_E = 40 N_fgk + 35 N_sab + 4(L-2) N_fgk, N counted by the instrumented oracle on the CPU "
                                "sample and scaled to the grid", counters=True))
    # C4: S(a,b)
    e_bins = synth.group_structure(70)
    G = len(e_bins) - 1
    for name, sab, tab in (("C4 H-in-H2O S(a,b) discrete (skewed), P5", synth.c4_sab("skewed"), False),
                           ("C4 H-in-H2O S(a,b) continuous, P5", synth.c4_sab("cont", n_eout=400), False),
                           ("C4 H-in-H2O S(a,b) discrete (skewed), 16-bin cosine histogram", synth.c4_sab("skewed"), True)):
        order = 16 if tab else 5
        L = order if tab else order + 1
        ds = scatt.DeviceSab(sab, ctx)
        Ein, _ = ds.egrid(e_bins)            # sab_egrid (src/sab.F90:460-568) on the device
        for _ in range(2):
            ctx.stats(reset=True)
            t0 = time.perf_counter()
            got = ds.calc(e_bins, ace.SCATT_TYPE_TABULAR if tab else ace.SCATT_TYPE_LEGENDRE, order, Ein)
            wall = time.perf_counter() - t0
        st = ctx.stats(reset=True)
        ds.clear()
        row = {"config": name, "evals": int(got.size), "evals_per_s_e2e": got.size / wall, "kernel_ms": st["kernel_ms"],
               "evals_per_s_kernel": got.size / (st["kernel_ms"] * 1e-3), "launches": int(st["launches"]), "NE": len(Ein)}
        n_mu = sab.n_inelastic_mu
        if sab.secondary_mode == ace.SAB_SECONDARY_CONT:
            f1 = sum((len(d.e_out) + 2 * G) * n_mu * (4 * (L - 2) + 2 * L) for d in sab.inelastic_data)
            by = 24.0 * G * L * len(Ein)
            row["roofline"] = {"bound": "hbm", "algorithmic_bytes": by, "achieved": by / (st["kernel_ms"] * 1e-3) / 1e9,
                               "peak": hbm_peak, "unit": "GB/s", "frac": by / (st["kernel_ms"] * 1e-3) / 1e9 / hbm_peak if hbm_peak else None,
                               "formula": "stage 2: 24 G L bytes per E_in (stage 1: %.3g flop once)" % f1,
                               "note": "launch-bound at this size: 5 kernels over %d E_in" % len(Ein)}
        else:
            f = float(len(Ein) * sab.n_inelastic_e_out * (17 + n_mu * (3 + 4 * (L - 2) + 2 * L)))
            row["roofline"] = {"bound": "fp64", "algorithmic_flops": f, "achieved": f / (st["kernel_ms"] * 1e-3) / 1e12,
                               "peak": peak_tflops, "unit": "TFLOP/s", "frac": f / (st["kernel_ms"] * 1e-3) / 1e12 / peak_tflops,
                               "formula": "F_F = NEo (17 + n_mu (3 + 4(L-2) + 2L)) per E_in",
                               "note": "launch-bound at this size: 4 kernels over %d E_in" % len(Ein)}
        if sample_cpu:
            from oracle import pyoracle
            idx = np.arange(len(Ein))[:: max(1, len(Ein) // 400)]
            t0 = time.perf_counter()
            ref = pyoracle.sab_calc(sab, e_bins, order, Ein[idx], tabular=tab)
            row["cpu"] = {"evals_per_s": ref.size / (time.perf_counter() - t0), "cores": 1, "kind": "port",
                          "sample": f"{len(idx)} evenly spread E_in"}
        rows.append(row)
    return rows


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
