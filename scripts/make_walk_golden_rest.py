#!/usr/bin/env python
"""Writes tests/golden/walk_vectors_rest.npz: results of the independent numpy evaluations of tests/walks.py for the
parity-unpinned routines that scripts/make_walk_golden.py does not cover -- S(a,b) elastic / discrete / continuous +
combine_sab_grid, law 9, integrate_file6_lab_leg, thin_grid, apply_tol_scatt.  Neither the oracle nor CUDA is used.
tests/test_oracle_golden.py holds the oracle against the file, tests/test_gpu_parity.py the CUDA path."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tests import walks  # noqa: E402


def build():
    out = {}
    for mode in ("equal", "skewed"):
        sab, e_bins, E = walks.sab_discrete_case(mode)
        out[f"sab_disc_{mode}_Ein"] = E
        out[f"sab_disc_{mode}"] = walks.walk_sab_discrete(sab, e_bins, E, mode)
    for elastic in ("coherent", "incoherent"):
        sab, e_bins, E = walks.sab_elastic_case(elastic)
        el, sig = walks.walk_sab_elastic(sab, e_bins, E, elastic)
        out[f"sab_el_{elastic}_Ein"], out[f"sab_el_{elastic}"], out[f"sab_el_{elastic}_sig"] = E, el, sig
    sab, e_bins, E = walks.sab_continuous_case()
    inel, comb = walks.walk_sab_continuous(sab, e_bins, E)
    out.update(sab_cont_Ein=E, sab_cont_inel=inel, sab_cont=comb)
    nuc, e_bins, params, spec, Ein = walks.law9_case()
    out.update(law9_Ein=Ein, law9_p0=np.array([walks.walk_law9(e_bins, spec, float(e)) for e in Ein]))
    p0, ratio = walks.walk_file6_lab()
    out.update(file6_lab_p0=p0, file6_lab_p1_over_p0=np.array(ratio))
    x, y, tokeep, tol = walks.thin_case()
    out.update(thin_keep=walks.walk_thin_grid(x, y, tokeep, tol))
    d, tol = walks.tol_case()
    out.update(tol_out=walks.walk_apply_tol(d, tol))
    return out


def main():
    path = os.path.join(ROOT, "tests", "golden", "walk_vectors_rest.npz")
    np.savez_compressed(path, **build())
    print("wrote", path)


if __name__ == "__main__":
    main()
