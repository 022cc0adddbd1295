#!/usr/bin/env python
"""Writes the fixtures under tests/golden/.

  reference_kats.json   the known-answer vectors the reference's own test program holds for this path
                        (transcribed, with their file:line; the Fortran itself cannot be run here -- no
                        Fortran compiler in the image)
  chi_vectors.npz       oracle outputs of the fission-spectrum integration (calc_chi) on the two synthetic
                        fissionable nuclides, merged grid and a dense grid
  oracle_vectors.npz    outputs of the pinned CPU oracle (oracle/, checked against the KATs by
                        tests/test_oracle_golden.py) on small seeded inputs of every integrator, for the
                        GPU parity tests and as a guard against the oracle drifting

Inputs are regenerated from ndpp_b200.synth with fixed seeds, so only outputs are stored.
Run from the repo root: python scripts/make_golden.py
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")

KATS = {
    "source": "ndpp/ndpp tests/test_scatt/test_scattdata.F90 and the Sage worksheets beside it (SURVEY.md 8c)",
    "G1_convert_file4_equiprobable_linear": {
        "cite": "tests/test_scatt/test_scattdata.F90:595-603 (34-value table), exact /= comparison",
        "mu": [-1.0, -0.5, 0.0, 0.5, 1.0],
        "distro": [8.8388347646636875e-02, 0.21338834765811932, 0.48385358672217688, 0.73943449322968258,
                   0.99212549203273326]},
    "G2_convert_file6_law44": {
        "cite": "tests/test_scatt/test_scattdata.F90:894-981, tolerance 1e-10",
        "mu": [-1.0, -0.5, 0.0, 0.5, 1.0],
        "R1_A1": [0.1565176427, 0.2580539668, 0.4254590641, 0.7014634088, 1.1565176427],
        "R0_A0p5": [0.5409883534, 0.4948293954, 0.4797586878, 0.4948293954, 0.5409883534]},
    "G4_mu_bounds_tolab": {
        "cite": "tests/test_scatt/test_scattdata.F90:1512-1569 (awr 0.999167, Q 0)",
        "cases": [[1.5, 1.0, 0.81666661634070423168], [20.0, 1.0, 0.22537631014397342822],
                  [20.0, 2.0, 0.31741314579775205019]]},
    "G5_int_pn_tablelin": {
        "cite": "tests/test_scatt/test_scattdata.F90:1650,1687-1692 + integrate_file4_leg_reference.sws: "
                "integral of 0.5(x+1) P_l over [-1,-0.75], l = 0..5",
        "values": [0.015625, -0.0130208333333333, 0.008544921875, -0.00341796875, -0.00105031331380208,
                   0.00387191772460938]},
    "G6_file6_lab_single_eout": {
        "cite": "tests/test_scatt/test_scattdata.F90:1762-1802", "group_1_2": [1.0, 1.0 / 3.0, 0.0, 0.0, 0.0, 0.0]},
    "G7_isotropic_cm_A2": {
        "cite": "tests/test_scatt/test_interp_distro.sws cell 24; test_scattdata.F90:2041-2042,2109-2110",
        "P1_over_P0": 1.0 / 3.0, "P2_over_P0": 0.0519541, "P3": 0.0, "P4_over_P0": -0.00115050, "tol": 1e-6},
}


def vectors():
    from ndpp_b200 import ace, egrid, synth
    from oracle import pyoracle
    from tests.util import small_heavy
    out = {}
    # C1 fixture (MT 51/52 labels)
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(13)
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    out["c1_Ein"] = Ein
    out["c1_elastic"] = rn.elastic(Ein)
    out["c1_inelastic"], out["c1_nu_inelastic"] = rn.inelastic(Ein)
    rn.close()
    # scaled-down heavy nuclide: levels (file 4 CM) + Law 44 continuum (unit base + file 6 CM)
    nuc = small_heavy()
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7)
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != ace.ELASTIC)
    Eel = nuc.energy[::16]
    Ein = nuc.energy[nuc.energy >= thr][::9]
    out["heavy_Eel"], out["heavy_Ein"] = Eel, Ein
    out["heavy_elastic"] = rn.elastic(Eel, n_threads=os.cpu_count())
    out["heavy_inelastic"] = rn.inelastic(Ein, n_threads=os.cpu_count())[0]
    rn.close()
    # H-1 free gas, a handful of E_in
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=1000)
    Ein = Ein[[5, 400, 700, 950]]
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    out["freegas_Ein"] = Ein
    out["freegas_elastic"] = rn.elastic(Ein, n_threads=os.cpu_count())
    rn.close()
    # S(a,b): discrete skewed, continuous
    e_bins = synth.group_structure(70)
    for mode, kw in (("skewed", {}), ("cont", {"elastic": "incoherent"})):
        sab = synth.c4_sab(mode, **kw)
        E = egrid.sab_egrid(sab, e_bins)[::160]
        out[f"sab_{mode}_Ein"] = E
        out[f"sab_{mode}"] = pyoracle.sab_calc(sab, e_bins, 5, E)
    # Legendre leaf
    rng = np.random.default_rng(20261018)
    xl = rng.uniform(-1, 0.98, 64); xh = xl + rng.uniform(1e-4, 0.02, 64)
    fl, fh = rng.uniform(0, 2, 64), rng.uniform(0, 2, 64)
    out["leaf_args"] = np.stack([xl, xh, fl, fh])
    out["leaf_integrals"] = np.stack([pyoracle.calc_int_pn_tablelin(8, *a) for a in zip(xl, xh, fl, fh)])
    return out


def chi_vectors():
    from ndpp_b200 import synth
    from oracle import pyoracle
    out = {}
    e_bins = synth.group_structure(70)
    dense = np.geomspace(1e-11, 20.0, 97)
    for name, mk in (("total", synth.fissile_total), ("partial", synth.fissile_partial)):
        for gname, grid in (("merged", None), ("dense", dense)):
            E, t, p, d = pyoracle.calc_chi(mk(), e_bins, E_grid=grid)
            out[f"{name}_{gname}_E"], out[f"{name}_{gname}_total"] = E, t
            out[f"{name}_{gname}_prompt"], out[f"{name}_{gname}_delay"] = p, d
    return out


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    cv = chi_vectors()
    np.savez_compressed(os.path.join(GOLD, "chi_vectors.npz"), **cv)
    if "--chi-only" in sys.argv:
        print({k: getattr(a, "shape", None) for k, a in cv.items()})
        sys.exit(0)
    json.dump(KATS, open(os.path.join(GOLD, "reference_kats.json"), "w"), indent=1)
    v = vectors()
    np.savez_compressed(os.path.join(GOLD, "oracle_vectors.npz"), **v)
    print({k: getattr(a, "shape", None) for k, a in v.items()})
    print(os.path.getsize(os.path.join(GOLD, "oracle_vectors.npz")), "bytes")
