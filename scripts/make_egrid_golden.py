#!/usr/bin/env python
"""Writes tests/golden/egrid_vectors.npz: incoming-energy grids of create_Ein_grid (src/scatt.F90:166-536) and sab_egrid
(src/sab.F90:460-568) from the independent numpy statement of the Fortran text (ndpp_b200/egrid.py: sorted unions,
math.log / math.exp of the C library) on glibc 2.39 -- the vectors the oracle's literal chain of merges and the CUDA path
are held against directly (tests/test_egrid.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ndpp_b200 import egrid, synth  # noqa: E402


def cases():
    eb = synth.group_structure(70)
    nuc = synth.heavy_nuclide(n_grid=300, n_levels=4, seed=7, n_ein_cont=6, np_cont=10, n_el_adist=8, n_lvl_adist=4, np_lvl=9)
    el, inel = egrid.create_Ein_grid(nuc, eb)
    out = {"heavy4_el": el, "heavy4_inel": inel}
    el, inel = egrid.create_Ein_grid(nuc, eb, extend_pts=7, inel_extend_pts=4)
    out.update({"heavy4_el_7_4": el, "heavy4_inel_7_4": inel})
    nuc3, eb3, _, _ = synth.c3_h1_freegas()
    out["h1_el"] = egrid.create_Ein_grid(nuc3, eb3)[0]
    out["sab_skewed_coherent"] = egrid.sab_egrid(synth.c4_sab(mode="skewed", elastic="coherent", n_ein=20, n_eout=12), eb)
    out["sab_cont_0"] = egrid.sab_egrid(synth.c4_sab(mode="cont"), eb, sab_epts_per_bin=0)
    return out


if __name__ == "__main__":
    path = os.path.join(ROOT, "tests", "golden", "egrid_vectors.npz")
    np.savez_compressed(path, **cases())
    print("wrote", path, os.path.getsize(path), "bytes")
