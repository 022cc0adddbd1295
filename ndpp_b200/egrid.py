"""Incoming-energy grid builders, host restatement (numpy).

In the reference these routines run in Fortran *before* the hot path and hand it the E_in grids:
  merge               src/array_merge.F90:13-107
  create_Ein_grid     src/scatt.F90:166-236   (combine_Eins :246, add_elastic_Eins :311,
                      add_one_more_point :426, add_inelastic_Eins :456)
  sab_egrid           src/sab.F90:460-568
The product builds them on the device (csrc/kernels_egrid.cuh: DeviceNuclide.create_ein_grid, DeviceSab.egrid).  This
module is the independent statement of the same text that holds the oracle's literal chain of merges
(oracle/egrid_ref.c) and the device path, bit for bit, in tests/test_egrid.py; the synthetic benchmark workloads
(synth.py) and chi.py / output.py use its `merge` and `binary_search` for host-side bookkeeping.  `merge` is a sorted
union, which equals the reference's two-pointer merge whenever neither input holds repeated values (the reference
keeps some repeats that occur inside one array); log and exp are the C library's, as the gfortran build calls them.
"""
from __future__ import annotations

import math

import numpy as np

from .ace import ELASTIC, SAB_SECONDARY_CONT, Nuclide, SAlphaBeta, iter_slots

MIN_EIN = 1e-14          # src/constants.F90:109
EXTEND_PTS = 50          # src/constants.F90:81
INEL_EXTEND_PTS = 30     # src/constants.F90:83


def _exp(x) -> np.ndarray:
    """exp / log of the C library, as the gfortran build calls them (numpy's SIMD loops differ from glibc in the last
    bit for a few per cent of the arguments, which would move grid points)."""
    return np.array([math.exp(v) for v in np.atleast_1d(x)])


def _log(x: float) -> float:
    return math.log(x) if x > 0.0 else float(np.log(x))


def merge(a, b) -> np.ndarray:
    a, b = np.asarray(a, float), np.asarray(b, float)
    if a.size == 0:
        return b.copy()
    if b.size == 0:
        return a.copy()
    u = np.union1d(a, b)
    # a zero that meets a larger value is replaced by MIN_EIN; a zero present in both stays (:49-59)
    if u[0] == 0.0 and ((a[0] == 0.0) != (b[0] == 0.0)):
        u[0] = MIN_EIN
    return u


def binary_search(arr, val) -> int:
    """1-based lower index, as src/search.F90:21-71 (val == last -> n-1)."""
    arr = np.asarray(arr)
    if val < arr[0] or val > arr[-1]:
        raise ValueError("Value outside of array during binary search")
    i = int(np.searchsorted(arr, val, side="right"))
    return min(i, len(arr) - 1)


def slot_summary(nuc: Nuclide, e_bins):
    """(is_init, MT, Q, E_grid) per ScattData slot: the part of scatt_init
    (src/scattdata_header.F90:78-271) the grid builders look at."""
    out = []
    has_adist = {}
    for idx, rxn, ed in iter_slots(nuc):
        MT = rxn.MT
        valid = (MT == ELASTIC or 11 <= MT <= 91) and MT not in (18, 19, 20, 21, 38)
        if valid and ed is not None and ed.law not in (3, 44, 61, 9, 4):
            valid = False
        if not valid:
            out.append((False, MT, rxn.Q_value, None))
            continue
        adist = has_adist.get(idx, rxn.adist is not None)
        iso = np.array([max(nuc.energy[rxn.threshold - 1], e_bins[0]), e_bins[-1]])
        if adist:
            use_edist = ed is not None and ed.law != 3
            egrid_ad = rxn.adist.energy if rxn.adist is not None else iso
        elif ed is not None:
            if ed.law in (4, 3, 9):
                use_edist = ed.law in (9, 4)
                has_adist[idx] = True
                egrid_ad = iso
            else:
                use_edist = True
                egrid_ad = None
        else:
            use_edist = False
            has_adist[idx] = True
            egrid_ad = iso
        if use_edist:
            d = np.asarray(ed.data)
            NR = int(d[0])
            NE = int(d[1 + 2 * NR])
            eg = d[2 + 2 * NR:2 + 2 * NR + NE].copy()
        else:
            eg = np.asarray(egrid_ad, float)
        out.append((True, MT, rxn.Q_value, eg))
    return out


def add_one_more_point(Ein):
    # the literal 1.0E-3 is single precision, promoted to double before the sum (:438)
    return np.concatenate([Ein, [Ein[-1] * (1.0 + float(np.float32(1.0e-3)))]])


def add_elastic_Eins(awr, kT, cutoff, E_bins, Ein, extend_pts=EXTEND_PTS):
    alpha = ((awr - 1.0) / (awr + 1.0)) ** 2
    lo_shift = 2.0 * kT * (awr + 1.0) / awr
    if cutoff != 0.0:
        new = []
        for g in range(len(E_bins) - 1):
            Ehi, Elo = E_bins[g + 1], E_bins[g]
            if Ehi <= cutoff:
                dElo = (_log(Ehi / (Ehi - lo_shift)) if Ehi - lo_shift > Elo else _log(Ehi / 1e-11)) / extend_pts
                pts = Ehi * _exp(np.arange(-extend_pts, 0) * dElo)
                new.append(pts[pts >= Elo])
            elif Elo < cutoff:
                dElo = _log(cutoff / (cutoff - lo_shift)) / extend_pts
                pts = cutoff * _exp(np.arange(-extend_pts, 0) * dElo)
                new.append(pts[pts > Elo])
        if new:
            Ein = merge(np.sort(np.concatenate(new)), Ein)
    dEhi = 7.0 * _log(1.0 / alpha) / extend_pts if alpha > 0 else np.inf
    new = []
    for g in range(len(E_bins) - 1):
        if E_bins[g] == 0.0:
            continue
        pts = E_bins[g] * _exp(np.arange(1, extend_pts) * dEhi)
        keep = pts < E_bins[g + 1]
        n = int(np.argmin(keep)) if not keep.all() else len(pts)  # the reference exits at the first failure
        new.append(pts[:n])
    if new:
        allp = np.sort(np.concatenate(new))
        if allp.size:
            Ein = merge(allp, Ein)
    return Ein


def add_inelastic_Eins(slots, awr, E_bins, thresh, Ein, inel_extend_pts=INEL_EXTEND_PTS):
    for is_init, MT, Qv, _ in slots:
        if not is_init:
            continue
        Q = -Qv
        if Q == 0.0:
            continue
        new = []
        for g in range(1, len(E_bins) - 1):  # g = 2 .. size-1 (1-based)
            Eg = E_bins[g]
            Ef = (1.0 + awr) / awr * Eg
            D = ((awr * awr) * (1.0 + Ef / Q) - 1.0) * (Ef / Q)
            with np.errstate(all="ignore"):
                sq = np.sqrt(D)
                Fp = (1.0 + sq) / (1.0 + Ef / Q)
                Fm = (1.0 - sq) / (1.0 + Ef / Q)
                Ecp = ((1.0 + awr) / awr * Q) / (1.0 - Fp * Fp / (awr * awr))
                Ecm = ((1.0 + awr) / awr * Q) / (1.0 - Fm * Fm / (awr * awr))
            Elo, Ehi = (Ecm, Ecp) if Ecp > Ecm else (Ecp, Ecm)
            Elo = max(Elo, thresh) if not np.isnan(Elo) else thresh
            Ehi = max(Ehi, thresh) if not np.isnan(Ehi) else thresh
            if Elo != Ehi:
                dE = _log(Ehi / Elo) / inel_extend_pts
                new.append(Elo * _exp(np.arange(1, inel_extend_pts) * dE))
        if new:
            Ein = merge(np.sort(np.concatenate(new)), Ein)
    return Ein


def create_Ein_grid(nuc: Nuclide, E_bins, extend_pts=EXTEND_PTS, inel_extend_pts=INEL_EXTEND_PTS):
    """Returns (Ein_el, Ein_inel); Ein_inel is None for a nuclide with elastic scattering only."""
    E_bins = np.asarray(E_bins, float)
    slots = slot_summary(nuc, E_bins)
    grid = np.asarray(nuc.energy, float)
    iEmax = len(grid) if E_bins[-1] >= grid[-1] else binary_search(grid, E_bins[-1])
    Ein_el = merge(grid[:iEmax], E_bins)
    # combine_Eins
    only_el = True
    new_grid = Ein_el[:1].copy()
    inel_thresh = E_bins[-1]
    for (is_init, MT, Q, eg), (idx, rxn, ed) in zip(slots, iter_slots(nuc)):
        if not is_init:
            continue
        if MT != ELASTIC:
            only_el = False
            inel_thresh = min(inel_thresh, nuc.energy[rxn.threshold - 1])
        if E_bins[0] >= eg[-1] or E_bins[-1] <= eg[0]:
            continue
        n = len(eg) if E_bins[-1] >= eg[-1] else binary_search(eg, E_bins[-1])
        new_grid = merge(eg[:n], new_grid)
    Ein_el = merge(new_grid, Ein_el)
    Ein_el = add_elastic_Eins(nuc.awr, nuc.kT, nuc.freegas_cutoff, E_bins, Ein_el, extend_pts)
    Ein_el = add_one_more_point(Ein_el)
    Ein_inel = None
    if not only_el:
        i = binary_search(Ein_el, inel_thresh)
        Ein_inel = Ein_el[i - 1:].copy()
        Ein_inel = add_inelastic_Eins(slots, nuc.awr, E_bins, inel_thresh, Ein_inel, inel_extend_pts)
        i = binary_search(Ein_inel, E_bins[-1])
        Ein_inel = add_one_more_point(Ein_inel[:i])
    return Ein_el, Ein_inel


def sab_egrid(sab: SAlphaBeta, energy_bins, sab_epts_per_bin=10, extend_pts=EXTEND_PTS):
    """src/sab.F90:460-568.  Note the reference expands every interval with EXTEND_PTS points whenever
    SAB_EPTS_PER_BIN is non-zero (:548-566)."""
    eb = np.asarray(energy_bins, float)
    if sab.elastic_e_in is not None:
        Ein = merge(merge(sab.inelastic_e_in, sab.elastic_e_in), eb)
        max_ein = max(sab.inelastic_e_in[-1], sab.elastic_e_in[-1])
    else:
        Ein = merge(sab.inelastic_e_in, eb)
        max_ein = sab.inelastic_e_in[-1]
    if sab.secondary_mode != SAB_SECONDARY_CONT:
        eo = np.asarray(sab.inelastic_e_out)        # [n_e_in][n_e_out]
        ei = np.asarray(sab.inelastic_e_in)
        new = []
        for i in range(len(ei) - 1):
            Eo1, Eo2 = eo[i], eo[i + 1]
            g1 = np.minimum(np.searchsorted(eb, Eo1, side="right"), len(eb) - 1)
            g2 = np.minimum(np.searchsorted(eb, Eo2, side="right"), len(eb) - 1)
            g2 = np.where(Eo2 < Eo1, g1, g2)        # :508-512 (the swap leaves g1 = g2)
            for j in np.nonzero(g2 > g1)[0]:
                g = np.arange(g1[j] + 1, g2[j] + 1)  # 1-based group-edge indices
                new.append((eb[g - 1] - Eo1[j]) / (Eo2[j] - Eo1[j]) * (ei[i + 1] - ei[i]) + ei[i])
        if new:
            Ein = merge(np.sort(np.concatenate(new)), Ein)
    i_max = binary_search(Ein, max_ein)
    base = Ein[:i_max]
    if sab_epts_per_bin == 0:
        return base.copy()
    out = []
    for k in range(i_max - 1):
        dE = _log(base[k + 1] / base[k]) / float(extend_pts + 1)
        seg = np.empty(extend_pts + 1)
        seg[0] = base[k]
        step = math.exp(dE)
        for q in range(1, extend_pts + 1):   # Ein(j) = Ein(j-1) * exp(dE), sequentially (:560-563)
            seg[q] = seg[q - 1] * step
        out.append(seg)
    out.append(base[-1:])
    return np.concatenate(out)
