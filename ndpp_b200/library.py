"""Library runs: many nuclides over the GPUs of one box (BASELINE.json configs[4]).

The reference distributes a library over MPI ranks as static contiguous blocks of *nuclides*
(src/ndpp.F90:941-948), which balances poorly: a U-238-shaped nuclide costs >1e3 x an H-1-shaped one.
Here the unit of work is (nuclide, matrix in {elastic, inelastic}, E_in tile); tiles are weighted with
the algorithmic-flop formulas of SURVEY 8d / DESIGN.md and dealt longest-processing-time-first to the
ranks.  Every (nuclide, E_in) column depends only on read-only tables, so there is no exchange step;
the one collective is the gather of the finished `[rows][G*L]` slabs to the rank that hands the
matrices back to the reference's driver for output (rank 0).  The top-of-grid copy rule
(src/scatt.F90:669,770) is applied after the gather because a tile's predecessor column may live on
another rank.

The planner needs only the shape of a nuclide (grid size, number of levels, thresholds, continuum
yes/no), so every rank derives the same plan without generating or parsing nuclides it does not own.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


@dataclass(frozen=True)
class NuclideShape:
    """What the cost model needs to know about a nuclide."""
    index: int
    n_el: int                 # elastic E_in points
    n_inel: int               # inelastic E_in points (estimate is fine; tiles are fractions)
    level_thresholds: Tuple[float, ...]   # MeV, discrete levels (file-4 CM integrations)
    cont_threshold: Optional[float]       # MeV, Law 44/61 continuum in the CM frame (file-6 CM), or None
    e_lo: float
    e_hi: float
    freegas_points: int = 0   # elastic E_in below the free-gas cutoff


@dataclass(frozen=True)
class WorkItem:
    nuclide: int
    matrix: str               # "el" | "inel"
    tile: int
    n_tiles: int
    cost: float               # algorithmic flops (model)


def flops_file4(G: int, L: int, M: int, g_act: int = 3) -> float:
    """F_A per (E_in, reaction), SURVEY 8d."""
    return 12 + 22 * G + 40 * g_act + (M + g_act) * (13 + 4 * (L - 2) + 10 * L)


def flops_file6_cm(G_b: int, L: int, M: int, K: int, NPu: int = 127) -> float:
    """F_B per (E_in, reaction), SURVEY 8d."""
    return 11.0 * M * NPu + G_b * K * (17 + 71.0 * M + (15 * L + 11) * (M - 1)) + 3 * G_b * L


def flops_freegas(G: int, L: int) -> float:
    """F_E per E_in with the evaluation counts measured on C3: ~3.2e7 fgk per E_in at L=4."""
    return 40 * 3.2e7 * (L / 4.0)


def tile_bounds(n: int, tile: int, n_tiles: int) -> Tuple[int, int]:
    """Rows [lo, hi) of tile `tile` of a grid of n points (first tiles take the remainder)."""
    base, rem = divmod(n, n_tiles)
    lo = tile * base + min(tile, rem)
    return lo, lo + base + (1 if tile < rem else 0)


def _log_grid_energy(shape: NuclideShape, frac: float) -> float:
    return float(np.exp(np.log(shape.e_lo) + frac * (np.log(shape.e_hi) - np.log(shape.e_lo))))


def make_items(shapes: Sequence[NuclideShape], G: int, L: int, M: int, K: int, tile_rows: int = 1024,
               world: Optional[int] = None) -> List[WorkItem]:
    """Work items of a library with their modelled cost.  E_in grids are taken as log-uniform between
    e_lo and e_hi for the purpose of the model (the synthetic libraries are; for real ACE grids the
    estimate only affects balance, never results).

    Continuum tiles are ~1e3 x heavier per row than the others.  With `world` given they are cut only as fine as
    balance needs -- the coarsest split (1, 2, 4, 8 x) whose heaviest item stays below 1/16 of a rank's share --
    because every extra tile is another launch of the persistent file-6 kernel with its own tail, and another
    rank that has to open the nuclide (300 nuclides on 8 GPUs: 4561 items / 445 opens instead of 8306 / 1269,
    modelled imbalance 1.003 instead of 1.015).  Without `world` the finest split is used."""
    if world is None:
        return _make_items(shapes, G, L, M, K, tile_rows, 8)
    for split in (1, 2, 4, 8):
        items = _make_items(shapes, G, L, M, K, tile_rows, split)
        total = sum(it.cost for it in items)
        if not items or max(it.cost for it in items) <= total / (16.0 * max(world, 1)):
            break
    return items


def _make_items(shapes, G, L, M, K, tile_rows, cont_split) -> List[WorkItem]:
    items: List[WorkItem] = []
    for s in shapes:
        nt = max(1, -(-s.n_el // tile_rows))
        for t in range(nt):
            lo, hi = tile_bounds(s.n_el, t, nt)
            cost = (hi - lo) * 2 * flops_file4(G, L, M)
            fg = max(0, min(hi, s.freegas_points) - lo)
            cost += fg * flops_freegas(G, L)
            items.append(WorkItem(s.index, "el", t, nt, float(cost)))
        if s.n_inel <= 0:
            continue
        thr = sorted(s.level_thresholds)
        # a nuclide whose inelastic slots are all (n,2n)-like (no level, no continuum threshold) starts at the grid
        e0 = min(list(thr) + ([s.cont_threshold] if s.cont_threshold is not None else []), default=s.e_lo)
        inel_lo = dict(e_lo=max(e0, s.e_lo), e_hi=s.e_hi)
        nt = max(1, -(-s.n_inel // tile_rows))
        if s.cont_threshold is not None:
            nt = max(nt, min(s.n_inel, cont_split * nt))
        for t in range(nt):
            lo, hi = tile_bounds(s.n_inel, t, nt)
            mid = (lo + hi) / 2.0 / max(s.n_inel, 1)
            E = float(np.exp(np.log(inel_lo["e_lo"]) + mid * (np.log(inel_lo["e_hi"]) - np.log(inel_lo["e_lo"]))))
            n_lev = sum(1 for x in thr if x < E)
            cost = (hi - lo) * 2 * n_lev * flops_file4(G, L, M)
            if s.cont_threshold is not None and E > s.cont_threshold:
                cost += (hi - lo) * flops_file6_cm(max(1, G - 5), L, M, K)
            items.append(WorkItem(s.index, "inel", t, nt, float(cost)))
    return items


def plan_lpt(items: Sequence[WorkItem], world: int, setup_cost: float = 5.0e10) -> List[List[WorkItem]]:
    """Longest-processing-time-first, aware of the per-rank cost of opening a nuclide (table upload +
    convert_distro, `setup_cost` in model flops): the heaviest item goes to the rank on which it would
    finish first, so light nuclides stay whole and only heavy ones are split.  Deterministic."""
    order = sorted(items, key=lambda it: (-it.cost, it.nuclide, it.matrix, it.tile))
    load = [0.0] * world
    have = [set() for _ in range(world)]
    out: List[List[WorkItem]] = [[] for _ in range(world)]
    for it in order:
        best = min(range(world), key=lambda r: (load[r] + it.cost + (0.0 if it.nuclide in have[r] else setup_cost), r))
        if it.nuclide not in have[best]:
            have[best].add(it.nuclide)
            load[best] += setup_cost
        load[best] += it.cost
        out[best].append(it)
    for r in range(world):
        out[r].sort(key=lambda it: (it.nuclide, it.matrix, it.tile))   # one table upload per nuclide
    return out


def plan_static_blocks(items: Sequence[WorkItem], shapes: Sequence[NuclideShape], world: int) -> List[List[WorkItem]]:
    """The reference's MPI partition (src/ndpp.F90:941-948): contiguous blocks of nuclides, first ranks
    take the remainder.  Kept for the load-balance comparison in scripts/run_library.py."""
    n = len(shapes)
    base, rem = divmod(n, world)
    owner: Dict[int, int] = {}
    k = 0
    for r in range(world):
        cnt = base + (1 if r < rem else 0)
        for s in shapes[k:k + cnt]:
            owner[s.index] = r
        k += cnt
    out: List[List[WorkItem]] = [[] for _ in range(world)]
    for it in items:
        out[owner[it.nuclide]].append(it)
    return out


def imbalance(plan: Sequence[Sequence[WorkItem]]) -> float:
    """max rank load / mean rank load of a plan (1.0 = perfect)."""
    loads = [sum(it.cost for it in p) for p in plan]
    mean = sum(loads) / max(len(loads), 1)
    return max(loads) / mean if mean > 0 else 1.0


def run_plan(my_items: Sequence[WorkItem], all_plans: Sequence[Sequence[WorkItem]],
             open_nuclide: Callable[[int], object], integrate: Callable[[object, WorkItem], "torch.Tensor"],
             close_nuclide: Callable[[object], None], GL: int, device, dst: int = 0, timers: Optional[dict] = None,
             rows_of: Optional[Callable[[WorkItem], int]] = None):
    """Integrate this rank's items and gather every slab to `dst` with one collective.

    open_nuclide(index) -> handle (uploads the tables once per nuclide on this rank);
    integrate(handle, item) -> `[rows][GL]` float64 tensor on `device`; or, when `rows_of(item)` (the tile's
    height, from the host-side grid) is given, integrate(handle, item, out) fills the `[rows][GL]` view `out`;
    returns on dst: {(nuclide, matrix): [(tile, n_tiles, tensor), ...]} with device tensors, else None.
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    import time
    tm = timers if timers is not None else {}
    tm.update(opens=0, open_s=0.0, integrate_s=0.0, pack_s=0.0, gather_s=0.0)

    def lap(key, t0):
        # host wall time of a phase, device drained first: the phases are separated for the report only
        if timers is not None:
            torch.cuda.synchronize(device)
        tm[key] += time.perf_counter() - t0

    # rows of every item, known to all ranks after one small all-reduce (tiles are fractions of grids
    # whose exact length only the owner knows)
    index = {}
    k = 0
    for r in range(world):
        for it in all_plans[r]:
            index[(r, it.nuclide, it.matrix, it.tile)] = k
            k += 1

    def share_rows(mine):
        rows = torch.zeros(max(k, 1), dtype=torch.int64, device=device)
        if len(mine):
            pos = torch.tensor([index[(rank, it.nuclide, it.matrix, it.tile)] for it in my_items], device=device)
            rows[pos] = torch.tensor(mine, dtype=torch.int64, device=device)
        if world > 1:
            dist.all_reduce(rows, op=dist.ReduceOp.SUM)
        rows_h = rows.tolist()
        per_rank = [sum(rows_h[index[(r, it.nuclide, it.matrix, it.tile)]] for it in all_plans[r]) for r in range(world)]
        return rows_h, max(max(per_rank), 1)

    def run_items(out_of):
        handle, cur = None, None
        res = []
        for j, it in enumerate(my_items):
            if it.nuclide != cur:
                if handle is not None:
                    close_nuclide(handle)
                t0 = time.perf_counter()
                handle, cur = open_nuclide(it.nuclide), it.nuclide
                tm["opens"] += 1
                tm["open_s"] += time.perf_counter() - t0      # uploads + convert_distro (synchronous)
            res.append(integrate(handle, it) if out_of is None else integrate(handle, it, out_of(j)))
        if handle is not None:
            close_nuclide(handle)
        return res

    t_all = time.perf_counter()
    if rows_of is not None:
        # The owner knows its tiles' heights from the host-side grids: one result buffer per rank is allocated
        # up front and every tile is integrated in place.  (Allocating a tensor per tile made the caching
        # allocator call cudaMalloc -- which synchronises the device -- thousands of times in a library run:
        # 13 s of a 35 s run at 300 nuclides on one GPU.)
        mine = [int(rows_of(it)) for it in my_items]
        rows_h, pad = share_rows(mine)
        flat = torch.empty((pad, GL), dtype=torch.float64, device=device)
        offs = np.concatenate([[0], np.cumsum(mine)]).astype(np.int64)
        run_items(lambda j: flat[int(offs[j]):int(offs[j + 1])])
        lap("integrate_s", t_all)
        tm["integrate_s"] -= tm["open_s"]
        t0 = time.perf_counter()
    else:
        slabs = run_items(None)
        lap("integrate_s", t_all)
        tm["integrate_s"] -= tm["open_s"]
        rows_h, pad = share_rows([s.shape[0] for s in slabs])
        t0 = time.perf_counter()
        flat = torch.zeros((pad, GL), dtype=torch.float64, device=device)
        o = 0
        for s in slabs:
            flat[o:o + s.shape[0]] = s
            o += s.shape[0]
    lap("pack_s", t0)
    t0 = time.perf_counter()
    if world > 1:
        parts = [torch.empty_like(flat) for _ in range(world)] if rank == dst else None
        dist.gather(flat, parts, dst=dst)
    else:
        parts = [flat]
    lap("gather_s", t0)
    if rank != dst:
        return None
    out: Dict[Tuple[int, str], list] = {}
    for r in range(world):
        o = 0
        for it in all_plans[r]:
            n = rows_h[index[(r, it.nuclide, it.matrix, it.tile)]]
            out.setdefault((it.nuclide, it.matrix), []).append((it.tile, it.n_tiles, parts[r][o:o + n]))
            o += n
    return out


def assemble(pieces, Ein, e_top: float):
    """Concatenate the tiles of one matrix in order and apply the top-of-grid copy rule."""
    import torch

    from .parallel import copy_top_columns
    pieces = sorted(pieces, key=lambda p: p[0])
    assert [p[0] for p in pieces] == list(range(pieces[0][1])), "missing tile"
    mat = torch.cat([p[2] for p in pieces], dim=0)
    assert mat.shape[0] == len(Ein), (mat.shape, len(Ein))
    return copy_top_columns(mat, torch.as_tensor(np.asarray(Ein), device=mat.device), e_top)
