"""Library runs: many nuclides over the GPUs of one box (BASELINE.json configs[4]).

The reference distributes a library over MPI ranks as static contiguous blocks of *nuclides*
(src/ndpp.F90:941-948), which balances poorly: a U-238-shaped nuclide costs >1e3 x an H-1-shaped one.
Here the unit of work is (nuclide, matrix in {elastic, inelastic}, E_in tile); tiles are weighted with the
algorithmic-flop formulas of SURVEY 8d and dealt longest-processing-time-first to the devices.  Planner,
per-device workers, the NCCL gather and the assembly all live in libndppgpu.so (csrc/group.cuh:
ndppgpu_plan_library, ndppgpu_library_run, ndppgpu_library_fetch); this module holds what the *host* driver
contributes -- the shapes of the nuclides, the callback that parses / builds a nuclide when a device asks for
it, and, when every GPU has its own process, the two small host-side exchanges (the NCCL id, the grid sizes)
that an MPI driver would do with MPI_Bcast / MPI_Allreduce and that run here over torch.distributed (any
backend: the CPU tests use gloo).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class NuclideShape:
    """What the cost model needs to know about a nuclide (ndppgpu_shape)."""
    index: int
    n_el: int                 # elastic E_in points
    n_inel: int               # inelastic E_in points (an estimate is fine for planning; see exchange_sizes)
    level_thresholds: Tuple[float, ...]   # MeV, discrete levels (file-4 CM integrations)
    cont_threshold: Optional[float]       # MeV, Law 44/61 continuum in the CM frame (file-6 CM), or None
    e_lo: float
    e_hi: float
    freegas_points: int = 0   # elastic E_in below the free-gas cutoff


def plan(shapes: Sequence[NuclideShape], G: int, L: int, M: int, K: int, world: int, tile_rows: int = 1024,
         setup_cost: float = 5.0e10, policy: str = "lpt"):
    """(items, modelled imbalance) from ndppgpu_plan_library; deterministic, so every rank derives the same plan
    from the shapes alone without generating or parsing nuclides it does not own."""
    from . import group
    return group.plan_library(shapes, G, L, M, K, world, tile_rows, setup_cost, policy)


def set_rows(items: List[dict], sizes: Dict[int, Tuple[int, int]]) -> List[dict]:
    """Tile heights from the true grid sizes {nuclide: (NE_el, NE_inel)} (the planner saw estimates)."""
    from . import group
    for it in items:
        lo, hi = group.tile_bounds(sizes[it["nuclide"]][it["matrix"]], it["tile"], it["n_tiles"])
        it["rows"] = hi - lo
    return items


# ---- host-side exchanges of the one-process-per-GPU form (MPI_Bcast / MPI_Allreduce in an MPI driver) ----------------
def broadcast_id(dist, device=None) -> Optional[bytes]:
    """The NCCL id of the group: made on rank 0 (ndppgpu_group_unique_id), broadcast to every rank."""
    import torch

    from .group import Group
    if dist.get_world_size() == 1:
        return None
    t = torch.zeros(128, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t = torch.tensor(list(Group.unique_id()), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    return bytes(t.cpu().tolist())


def exchange_sizes(dist, n_nuclides: int, mine: Dict[int, Tuple[int, int]], device=None) -> Dict[int, Tuple[int, int]]:
    """Every rank learns the grid sizes of every nuclide: each contributes the ones it has parsed."""
    import torch
    t = torch.zeros((n_nuclides, 2), dtype=torch.int64, device=device)
    for i, (a, b) in mine.items():
        t[i, 0], t[i, 1] = int(a), int(b)
    if dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    h = t.cpu().tolist()
    return {i: (int(h[i][0]), int(h[i][1])) for i in range(n_nuclides)}


def owners(items: Sequence[dict]) -> Dict[int, set]:
    """{rank: nuclides it opens}"""
    out: Dict[int, set] = {}
    for it in items:
        out.setdefault(it["rank"], set()).add(it["nuclide"])
    return out


# ---- C5: the synthetic 300-nuclide library --------------------------------------------------------------------------------
def run_c5(group, n_nuclides: int = 300, dist=None, policy: str = "lpt", tile_rows: int = 1024, ne_hi: int = 40000,
           keep: Optional[Callable] = None, device=None) -> Optional[dict]:
    """BASELINE.json configs[4]: `n_nuclides` of the 300 synthetic nuclides (examples/ndpp.xml shape: P5, 70 groups,
    mu_bins 2001) on a device group.  `dist` is torch.distributed when every GPU has its own process (then every rank
    calls this), None when the group spans the GPUs of this process.  Returns the report on the root process.

    Untimed: synthetic generation of the nuclides a process owns (stands for ACE parsing, which is outside the path).
    Timed (host wall clock around ndppgpu_library_run, max over ranks): table uploads, convert_distro, integration of every
    tile, the NCCL gather to the root device."""
    import time

    from . import ace, scatt, synth
    from .group import LibraryRun

    world, first, n_local = group.world, group.first, group.n_local
    specs = synth.c5_library(300, ne_hi=ne_hi)[:n_nuclides]
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, mu_bins=2001)
    G, L, M, K = 70, 6, params.mu_bins, params.ne_per_grp
    shapes = [synth.c5_shape(s) for s in specs]
    items, imb = plan(shapes, G, L, M, K, world, tile_rows, policy=policy)
    _, imb_static = plan(shapes, G, L, M, K, world, tile_rows, policy="static")
    own = owners(items)
    mine = sorted(set().union(*[own.get(r, set()) for r in range(first, first + n_local)]))
    parsed = {i: synth.c5_nuclide(specs[i]) for i in mine}
    sizes = {i: (len(p[1]), 0 if p[2] is None else len(p[2])) for i, p in parsed.items()}
    if dist is not None:
        sizes = exchange_sizes(dist, len(specs), sizes, device)
    set_rows(items, sizes)

    def open_nuclide(i, ctx):
        nuc, Eel, Einel = parsed[i]
        return scatt.DeviceNuclide(nuc, e_bins, params, ctx), Eel, (Einel if Einel is not None else np.zeros(0))

    # warm-up, untimed: one small nuclide of the heaviest shape on every local device (lazy kernel loading, growth of the
    # stream-ordered memory pool); NCCL sets its point-to-point connections up on first use, so a tiny library run with
    # one item per device goes first as well
    wn = synth.heavy_nuclide(n_grid=1500, n_levels=4, seed=99)
    for li in range(n_local):
        wd = scatt.DeviceNuclide(wn, e_bins, params, group.ctx(li))
        wd.elastic(wn.energy[:64])
        wd.inelastic(wn.energy[-64:])
        wd.clear()
    w_items = [dict(nuclide=0, matrix=0, tile=r, n_tiles=world, rank=r, rows=0, cost=1.0) for r in range(world)]
    set_rows(w_items, {0: (len(wn.energy), 0)})
    wr = LibraryRun(group, G, L, False, w_items)
    wr.run(lambda i, ctx: (scatt.DeviceNuclide(wn, e_bins, params, ctx), wn.energy, np.zeros(0)))
    wr.close()
    for li in range(n_local):
        group.ctx(li).stats(reset=True)
    group.gathered_bytes(reset=True)

    run = LibraryRun(group, G, L, False, items)
    if dist is not None:
        dist.barrier()
    t0 = time.perf_counter()
    rep = run.run(open_nuclide)
    wall = time.perf_counter() - t0
    k_max, k_sum, dev_s = rep["kernel_s_max"], rep["kernel_s_sum"], rep["device_s_max"]
    if dist is not None:
        import torch
        t = torch.tensor([wall, k_max, dev_s], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        s = torch.tensor([k_sum], dtype=torch.float64, device=device)
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        wall, k_max, dev_s, k_sum = float(t[0]), float(t[1]), float(t[2]), float(s[0])
    out = None
    if group.is_root:
        out = {"config": f"C5 library: {len(specs)} of 300 synthetic nuclides, P5, 70 groups, mu_bins 2001",
               "n_gpus": world, "processes": 1 if dist is None else world, "plan": policy, "work_items": len(items),
               "nuclide_opens": sum(len(v) for v in own.values()), "moment_evals": int(rep["moment_evals"]),
               "device_s": dev_s, "wall_s": wall, "evals_per_s": rep["moment_evals"] / dev_s,
               "timing": "device_s = CUDA events on each device from its first upload to the end of its part of the "
                         "gather, max over devices; wall_s = host clock around ndppgpu_library_run, max over ranks",
               "kernel_s_slowest_device": k_max,
               "measured_imbalance": k_max / (k_sum / world) if k_sum > 0 else None,
               "model_imbalance": {"lpt": imb if policy == "lpt" else plan(shapes, G, L, M, K, world, tile_rows)[1],
                                   "static": imb_static},
               "root_process": {k: rep[k] for k in ("compute_s", "gather_s", "open_s_max", "integrate_s_max", "alloc_s_max")},
               "gathered_bytes": group.gathered_bytes()}
        if keep is not None:    # e.g. the parity check of tests/util.py: the matrices are still on the root device
            keep(out, run, specs, parsed, e_bins, params)
    run.close()
    return out
