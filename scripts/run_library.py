#!/usr/bin/env python
"""C5: a synthetic library (examples/ndpp.xml shape, P5, 70 groups) over the GPUs of one box.

Launch with torchrun (one rank per GPU) or plainly for one GPU.  Work items are (nuclide, matrix, E_in
tile), cost-weighted and dealt longest-processing-time-first (ndpp_b200/library.py); the finished slabs
are gathered to rank 0 with one NCCL collective and assembled there.  Prints one JSON line on rank 0:
whole-job moment evaluations per second (CUDA-event time of the slowest rank + gather), the modelled
imbalance of the LPT plan next to the reference's static nuclide blocks (src/ndpp.F90:941-948), and a
parity check of a sampled nuclide against the CPU oracle.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    from ndpp_b200 import ace, library, scatt, synth
    from ndpp_b200.capi import Context

    ap = argparse.ArgumentParser()
    ap.add_argument("--nuclides", type=int, default=24)
    ap.add_argument("--tile-rows", type=int, default=1024)
    ap.add_argument("--plan", default="lpt", choices=["lpt", "static"])
    ap.add_argument("--check", type=int, default=1, help="nuclides checked against the CPU oracle on rank 0")
    ap.add_argument("--ne-hi", type=int, default=40000)
    ap.add_argument("--phases", action="store_true",
                    help="report host wall time per phase (opens / integrate / pack / gather); drains the device "
                         "between phases, so use a run without it for the headline number")
    args = ap.parse_args()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ctx = Context(local)
    lib_stream = torch.cuda.ExternalStream(ctx.stream, device=dev)

    specs = synth.c5_library(300, ne_hi=args.ne_hi)[:args.nuclides]
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, mu_bins=2001)
    G, L, M, K = 70, 6, params.mu_bins, params.ne_per_grp
    GL = G * L
    shapes = [synth.c5_shape(s) for s in specs]
    items = library.make_items(shapes, G, L, M, K, tile_rows=args.tile_rows, world=world)
    plans = {"lpt": library.plan_lpt(items, world), "static": library.plan_static_blocks(items, shapes, world)}
    plan = plans[args.plan]
    spec_of = {s[0]: s for s in specs}
    grids = {}

    # host-side state of the reference pipeline: the parsed nuclides of this rank (ACE parsing stays on the
    # host and is outside the path; here: synthetic generation, untimed)
    parsed = {i: synth.c5_nuclide(spec_of[i]) for i in sorted({it.nuclide for it in plan[rank]})}

    def open_nuclide(i):
        nuc, Eel, Einel = parsed[i]
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        Eel_d = torch.from_numpy(Eel).to(dev)
        Ein_d = torch.from_numpy(Einel).to(dev) if Einel is not None else None
        grids[i] = (Eel, Einel)
        return (dn, Eel_d, Ein_d)

    def rows_of(it):
        _, Eel, Einel = parsed[it.nuclide]
        lo, hi = library.tile_bounds(len(Eel if it.matrix == "el" else Einel), it.tile, it.n_tiles)
        return hi - lo

    def integrate(h, it, out):
        dn, Eel_d, Ein_d = h
        E = Eel_d if it.matrix == "el" else Ein_d
        lo, hi = library.tile_bounds(E.numel(), it.tile, it.n_tiles)
        assert out.shape[0] == hi - lo
        if hi > lo:
            if it.matrix == "el":
                dn.elastic_dev(E[lo:hi], out)
            else:
                dn.inelastic_dev(E[lo:hi], out)
        return out

    def close_nuclide(h):
        h[0].clear()

    # warm-up, untimed: one small nuclide of the heaviest shape through both matrices (lazy kernel loading,
    # growth of the stream-ordered memory pool, NCCL channel set-up)
    wn = synth.heavy_nuclide(n_grid=1500, n_levels=4, seed=99)
    wd = scatt.DeviceNuclide(wn, e_bins, params, ctx)
    wE = torch.from_numpy(wn.energy).to(dev)
    wo = torch.empty((wE.numel(), GL), dtype=torch.float64, device=dev)
    wd.elastic_dev(wE, wo)
    wd.inelastic_dev(wE, wo)
    wd.clear()
    if world > 1:
        # NCCL sets up its all-reduce rings and its point-to-point connections lazily, on the first collective of
        # each kind: one small all-reduce and one small gather before the clock starts
        dist.all_reduce(torch.zeros(1, device=dev))
        w_flat = torch.zeros((4, GL), dtype=torch.float64, device=dev)
        dist.gather(w_flat, [torch.empty_like(w_flat) for _ in range(world)] if rank == 0 else None, dst=0)
    ctx.stats(reset=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(lib_stream)
    with torch.cuda.stream(lib_stream):
        phases = {} if args.phases else None
        got = library.run_plan(plan[rank], plan, open_nuclide, integrate, close_nuclide, GL, dev, timers=phases,
                               rows_of=rows_of)
    e1.record(lib_stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    wall = time.perf_counter() - t0
    ms = e0.elapsed_time(e1)
    st = ctx.stats(reset=True)
    ph = phases or {}
    busy = torch.tensor([ms, st["kernel_ms"]] + [float(ph.get(k, 0.0)) for k in
                        ("opens", "open_s", "integrate_s", "pack_s", "gather_s")] +
                        [st["host_call_ms"], st["host_alloc_ms"], st["host_sync_ms"]], dtype=torch.float64, device=dev)
    if world > 1:
        all_busy = [torch.empty_like(busy) for _ in range(world)]
        dist.all_gather(all_busy, busy)
    else:
        all_busy = [busy]
    if rank == 0:
        evals = 0
        for (i, m), pieces in got.items():
            evals += sum(p[2].numel() for p in pieces)
        ms_ranks = [float(b[0]) for b in all_busy]
        line = {"config": f"C5 library: {len(specs)} of 300 synthetic nuclides, P5, 70 groups, mu_bins 2001",
                "n_gpus": world, "plan": args.plan, "work_items": len(items), "tile_rows": args.tile_rows,
                "moment_evals": int(evals), "seconds_wall": wall, "ms_device_max_over_ranks": max(ms_ranks),
                "ms_device_per_rank": ms_ranks, "ms_integrator_kernels_per_rank": [float(b[1]) for b in all_busy],
                "evals_per_s": evals / wall,
                "model_imbalance": {k: library.imbalance(v) for k, v in plans.items()},
                "measured_imbalance": max(float(b[1]) for b in all_busy) / (sum(float(b[1]) for b in all_busy) / world)}
        for j, k in enumerate(("host_call_ms", "host_alloc_ms", "host_sync_ms")):
            line[k + "_per_rank"] = [round(float(b[7 + j]), 1) for b in all_busy]
        if args.phases:
            for j, k in enumerate(("opens", "open_s", "integrate_s", "pack_s", "gather_s")):
                line["phase_" + k + "_per_rank"] = [round(float(b[2 + j]), 4) for b in all_busy]
        # parity of sampled nuclides (assembled on rank 0) against the oracle
        if args.check > 0:
            from oracle import pyoracle
            rng = np.random.default_rng(5)
            worst = {"cells": 0, "outside_floor": 0, "max_abs": 0.0}
            for i in [s[0] for s in specs][:: max(1, len(specs) // args.check)][:args.check]:
                nuc, Eel, Einel = synth.c5_nuclide(spec_of[i])
                rn = pyoracle.RefNuclide(nuc, e_bins, params)
                rn.convert_distro()
                for m, E in (("el", Eel), ("inel", Einel)):
                    if E is None or (i, m) not in got:
                        continue
                    mat = library.assemble(got[(i, m)], E, e_bins[-1]).cpu().numpy().reshape(len(E), G, L)
                    idx = np.sort(rng.choice(np.nonzero(E <= e_bins[-1])[0], min(24, len(E)), replace=False))
                    ref = rn.elastic(E[idx], n_threads=os.cpu_count()) if m == "el" else \
                        rn.inelastic(E[idx], n_threads=os.cpu_count())[0]
                    err = np.abs(mat[idx] - ref)
                    p0 = np.abs(ref[:, :, 0]).sum(axis=1)[:, None, None]
                    worst["cells"] += int(err.size)
                    worst["outside_floor"] += int((err > 1e-9 * np.abs(ref) + 1e-8 * p0 + 1e-12).sum())
                    worst["max_abs"] = max(worst["max_abs"], float(err.max()))
                rn.close()
            line["parity_vs_oracle"] = worst
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
