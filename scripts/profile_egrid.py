#!/usr/bin/env python
"""One create_Ein_grid of the C2 nuclide on the device (row N3), for the ncu launch list under profiles/:
   ncu --clock-control none --metrics gpu__time_duration.sum --csv --log-file gpurun_out/egrid_launches.csv python scripts/profile_egrid.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ndpp_b200 import scatt, synth  # noqa: E402

nuc, e_bins, params = synth.c2_u238()[:3]
dn = scatt.DeviceNuclide(nuc, e_bins, params)
for _ in range(2):
    (p_el, n_el), (p_in, n_in), status = dn.create_ein_grid(host=False)
print(n_el, n_in, status)
dn.clear()
