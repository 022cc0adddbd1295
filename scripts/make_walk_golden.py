#!/usr/bin/env python
"""Writes tests/golden/walk_vectors.npz: moments computed by the literal walks of the Fortran text (no oracle, no CUDA):

  freegas_*   tests/freegas_walk.py (pure-Python transcription of src/freegas.F90) at the reference's DEFAULT adaptive
              tolerances, H-1 with an isotropic CM table, 6 groups, P0..P2, two incoming energies (minutes of CPU time)
  file6_*     the numpy walk of unit-base interpolation + integrate_file6_cm_leg in tests/test_oracle_golden.py on the
              Law-44 continuum of tests.util.small_heavy(awr=236.0058), 24 groups, P0..P4, three incoming energies

The only oracle products used are the converted uniform-mu tables (pinned by the reference's own KATs G1 / G2).
tests/test_oracle_golden.py checks the oracle against these vectors, tests/test_gpu_parity.py the CUDA path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndpp_b200 import ace, synth  # noqa: E402
from oracle import pyoracle  # noqa: E402
from tests.freegas_walk import make_walk  # noqa: E402
from tests.test_oracle_golden import _walk_file6_cm_leg, _walk_unitbase  # noqa: E402
from tests.util import small_heavy  # noqa: E402


def freegas_case():
    kT = synth.KT_293K
    energy = np.geomspace(1e-11, 20.0, 100)
    nuc = ace.Nuclide(awr=0.999167, kT=kT, energy=energy, elastic=np.full(100, 20.0),
                      reactions=[ace.Reaction(MT=2, threshold=1)], freegas_cutoff=400 * kT)
    e_bins = np.array([0.0, 1e-9, 2e-8, 6e-8, 2e-7, 1e-6, 20.0])
    return nuc, e_bins, ace.Params(order=2, mu_bins=2001), np.array([0.7, 6.0]) * kT


def file6_case():
    nuc = small_heavy(awr=236.0058, first_level=0.0449, level_step=0.05)
    return nuc, synth.group_structure(24, 1e-4, 20.0), ace.Params(order=4, mu_bins=201)


def main():
    out = {}
    nuc, e_bins, params, Ein = freegas_case()
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    M = params.mu_bins
    gmu = list(-1.0 + np.arange(M) * (2.0 / (M - 1)))
    gmu[-1] = 1.0
    walk = make_walk(nuc.awr, nuc.kT, list(rn.get_table(0, 1)[0][:, 0]), gmu, 3, params.adaptive_mu_tol,
                     params.adaptive_mu_its, params.adaptive_eout_tol, params.adaptive_eout_its, params.sab_threshold,
                     params.brent_mu_thresh)
    res = []
    for E in Ein:
        m, n = walk(float(E), list(e_bins))
        print(f"free gas E = {E:.4e}: {n} kernel evaluations", flush=True)
        res.append(m)
    out.update(freegas_Ein=Ein, freegas_e_bins=e_bins, freegas_moments=np.array(res))

    nuc, e_bins, params = file6_case()
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    s = [k for k in range(rn.n_slots) if rn.slot_info(k)["is_init"] and rn.slot_info(k)["law"] == 44][0]
    eg = rn.slot_egrid(s)
    mu = -1.0 + 2.0 * np.arange(201) / 200.0
    mu[-1] = 1.0
    Ein = np.array([eg[0] * 1.01, 0.5 * (eg[2] + eg[3]), 0.93 * eg[-1]])
    res = []
    for E in Ein:
        iE = min(int(np.searchsorted(eg, E, side="right")), len(eg) - 1)
        rows = []
        for i in (iE, iE + 1):
            d, Eo, pdf, _, intt = rn.get_table(s, i)
            rows.append((d, Eo, pdf, intt))
        Eout, pdf, fEmu = _walk_unitbase(E, eg[iE - 1], rows[0], eg[iE], rows[1])
        res.append(_walk_file6_cm_leg(fEmu, mu, E, nuc.awr, Eout, pdf, e_bins, 5))
    out.update(file6_slot=s, file6_Ein=Ein, file6_moments=np.array(res))
    path = os.path.join(ROOT, "tests", "golden", "walk_vectors.npz")
    np.savez_compressed(path, **out)
    print("wrote", path)


if __name__ == "__main__":
    main()
