"""Host-side mirror of the reference's fission-spectrum orchestration, src/chi.F90.

`calc_chi` keeps the reference's name, argument meaning and results (src/chi.F90:21-163): it lists the ChiData
objects (prompt laws of every fission reaction with their nested laws, then one delayed law per precursor group,
:47-93), merges their incoming-energy grids (:96-112) and hands the E_in loop (:120-153) to `ndppgpu_chi`.
`print_chi` lives in output.py.  All numerical work happens in the CUDA library; there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np

from .ace import NU_NONE, DistEnergy, Nuclide
from .capi import Context, check, dp, f64
from .egrid import merge


class ChiSlotC(C.Structure):
    """ndppgpu_chi_slot (include/ndppgpu.h)."""
    _fields_ = [("law", C.c_int), ("delayed", C.c_int), ("precursor", C.c_int), ("threshold", C.c_int),
                ("use_pvalid", C.c_int), ("n_sigma", C.c_int), ("sigma_off", C.c_int), ("data_off", C.c_int),
                ("pvalid_off", C.c_int), ("reserved", C.c_int)]


def fission_xs(nuc: Nuclide) -> np.ndarray:
    """nuc % fission: the fission cross sections summed on the nuclide grid (src/ace.F90:823-828)."""
    if nuc.fission is not None:
        return f64(nuc.fission)
    fis = np.zeros(len(nuc.energy))
    for i in nuc.index_fission:
        r = nuc.reactions[i]
        fis[r.threshold - 1:r.threshold - 1 + len(r.sigma)] += r.sigma
    return fis


def chi_data(nuc: Nuclide) -> Tuple[List[dict], np.ndarray]:
    """The ChiData list of calc_chi (src/chi.F90:47-93) flattened for the C-ABI: per slot a dict of the
    ndppgpu_chi_slot fields plus `E_grid`, and the pool the offsets point into."""
    if not nuc.fissionable:
        raise ValueError("calc_chi is only called for fissionable nuclides (src/ndpp.F90:712)")
    fis = fission_xs(nuc)
    pool: List[np.ndarray] = []
    size = 0

    def put(a) -> int:
        nonlocal size
        a = f64(a)
        pool.append(a)
        size += len(a)
        return size - len(a)

    def grid_of(ed: DistEnergy) -> np.ndarray:
        NR = int(ed.data[0])                      # chi_init, src/chidata_header.F90:97-106
        if 1 + 2 * NR >= len(ed.data):
            raise ValueError(f"energy law {ed.law}: chi_init reads NR, NE and an incoming-energy grid from the law data "
                             "(src/chidata_header.F90:97-106); this law's data does not start with one")
        NE = int(ed.data[1 + 2 * NR])
        return f64(ed.data[2 + 2 * NR:2 + 2 * NR + NE])

    slots: List[dict] = []
    for i in nuc.index_fission:
        rxn = nuc.reactions[i]
        sigma = fis if rxn.MT == 18 else f64(rxn.sigma)          # :70-74
        s_off = put(sigma)
        ed: Optional[DistEnergy] = rxn.edist
        if ed is None:
            raise ValueError(f"fission reaction MT {rxn.MT} has no energy distribution")
        while ed is not None:
            use_pv = ed.next is not None and ed.p_valid is not None and len(ed.p_valid.nbt) > 0
            slots.append(dict(law=int(ed.law), delayed=0, precursor=0, threshold=int(rxn.threshold), use_pvalid=int(use_pv),
                              n_sigma=len(sigma), sigma_off=s_off, data_off=put(ed.data),
                              pvalid_off=put(ed.p_valid.flatten()) if use_pv else 0, E_grid=grid_of(ed)))
            ed = ed.next
    for k, ed in enumerate(nuc.nu_d_edist):
        slots.append(dict(law=int(ed.law), delayed=1, precursor=k + 1, threshold=0, use_pvalid=0, n_sigma=0, sigma_off=0,
                          data_off=put(ed.data), pvalid_off=0, E_grid=grid_of(ed)))
    return slots, (np.concatenate(pool) if pool else np.zeros(0))


def chi_grid(slots: List[dict]) -> np.ndarray:
    """The union grid of calc_chi, src/chi.F90:96-112 (array_merge.F90 merge, applied slot by slot)."""
    E = slots[0]["E_grid"].copy()
    for s in slots[1:]:
        E = merge(E, s["E_grid"])
    return f64(E)


def slots_c(slots: List[dict]):
    arr = (ChiSlotC * len(slots))()
    for a, s in zip(arr, slots):
        for k, _ in ChiSlotC._fields_:
            setattr(a, k, int(s.get(k, 0)))
    return arr


def calc_chi(nuc: Nuclide, E_bins, ctx: Optional[Context] = None, E_grid=None):
    """calc_chi(nuc, E_bins, E_grid, chi_total, chi_prompt, chi_delay), src/chi.F90:21.  Returns
    (E_grid[NE], chi_total[NE, G], chi_prompt[NE, G], chi_delay[n_precursor, NE, G]) -- the C views of the
    Fortran arrays chi_total(G, NE), chi_prompt(G, NE), chi_delay(G, NE, n_precursor)."""
    from .scatt import default_context
    ctx = ctx or default_context()
    if nuc.nu_t_type == NU_NONE:
        raise ValueError(f"No neutron emission data for table: {nuc.name}")    # src/fission.F90:29
    slots, pool = chi_data(nuc)
    E_grid = chi_grid(slots) if E_grid is None else f64(E_grid)
    E_bins = f64(E_bins)
    G, NE, n_prec = len(E_bins) - 1, len(E_grid), nuc.n_precursor
    energy, fis = f64(nuc.energy), fission_xs(nuc)
    nu_t = f64(nuc.nu_t_data)
    nu_d = f64(nuc.nu_d_data) if nuc.nu_d_data is not None else np.zeros(0)
    prec = f64(nuc.nu_d_precursor_data) if nuc.nu_d_precursor_data is not None else np.zeros(0)
    chi_t, chi_p, chi_d = np.empty((NE, G)), np.empty((NE, G)), np.empty((n_prec, NE, G))
    sc = slots_c(slots)
    check(ctx.lib.ndppgpu_chi(ctx.h, len(energy), dp(energy), dp(fis), int(nuc.nu_t_type), dp(nu_t), len(nu_t),
                              int(nuc.nu_d_type), dp(nu_d), len(nu_d), n_prec, dp(prec), len(prec), len(slots), sc,
                              dp(pool), len(pool), dp(E_bins), len(E_bins), dp(E_grid), NE, dp(chi_t), dp(chi_p),
                              dp(chi_d)), ctx.h)
    return E_grid, chi_t, chi_p, chi_d
