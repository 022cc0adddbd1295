#!/usr/bin/env python
"""One small invocation of a single path, for ncu captures (kept short: ncu replays every kernel)."""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndpp_b200 import ace, egrid, scatt, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--case", default="c3")
ap.add_argument("--n", type=int, default=40)
a = ap.parse_args()
if a.case == "c3":
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=a.n)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    out = dn.elastic(Ein)
    print("c3", out.shape, float(out[:, :, 0].sum()))
elif a.case == "c4":
    sab = synth.c4_sab("skewed")
    e_bins = synth.group_structure(70)
    out = scatt.DeviceSab(sab).calc(e_bins, 0, 5, egrid.sab_egrid(sab, e_bins))
    print("c4", out.shape)
elif a.case == "c2levels":
    nuc = synth.heavy_nuclide(n_grid=a.n * 50, with_continuum=False)
    e_bins = synth.group_structure(70)
    dn = scatt.DeviceNuclide(nuc, e_bins, ace.Params(order=7))
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != 2)
    out, _ = dn.inelastic(nuc.energy[nuc.energy >= thr])
    el = dn.elastic(nuc.energy)
    print("c2levels", out.shape, el.shape)
print(scatt.default_context().stats())
