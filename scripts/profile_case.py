#!/usr/bin/env python
"""One pass of one BASELINE configuration through the C-ABI, for profiling under ncu (never for timing):
    python scripts/profile_case.py --case c1|c2|c3|c4d|c4c [--n N]
c2: U-238 shape (--n = grid points, default 20000); c3: H-1 free gas (--n = E_in points, default 1000);
c4d / c4c: S(a,b) discrete / continuous; c1: the tests/test_scatt fixture."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    from ndpp_b200 import ace, egrid, scatt, synth
    ap = argparse.ArgumentParser()
    ap.add_argument("--case", required=True)
    ap.add_argument("--n", type=int, default=0)
    a = ap.parse_args()
    if a.case == "c1":
        nuc, e_bins, params = synth.c1_fixture()
        Ein = synth.c1_ein_grid(997)
        dn = scatt.DeviceNuclide(nuc, e_bins, params)
        dn.elastic(Ein), dn.inelastic(Ein)
    elif a.case == "c2":
        nuc, e_bins, params, Eel, Einel = synth.c2_u238(n_grid=a.n or 20000)
        dn = scatt.DeviceNuclide(nuc, e_bins, params)
        dn.elastic(Eel), dn.inelastic(Einel)
    elif a.case == "c3":
        nuc, e_bins, params, Ein = synth.c3_h1_freegas()
        dn = scatt.DeviceNuclide(nuc, e_bins, params)
        dn.elastic(Ein[:: max(1, len(Ein) // (a.n or len(Ein)))])
    elif a.case in ("c4d", "c4c"):
        sab = synth.c4_sab("skewed") if a.case == "c4d" else synth.c4_sab("cont", n_eout=400)
        e_bins = synth.group_structure(70)
        scatt.calc_scattsab(sab, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 2001, egrid.sab_egrid(sab, e_bins))
    else:
        raise SystemExit("unknown case")


if __name__ == "__main__":
    main()
