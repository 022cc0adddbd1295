#!/usr/bin/env python
"""Times the free-gas path (C3, 293.6 K, 1000 E_in; second pass) under different settings of the work-item environment
variables (NDPPGPU_FG_SPLIT / _SPLIT_LATE / _LATE_ITEMS): `python scripts/ab/fg_env.py "2 2 0" "0 0 0" "2 0 50000"`."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, hashlib
sys.path.insert(0, %r)
from ndpp_b200 import scatt, synth
ctx = scatt.default_context()
for kT, T in ((synth.KT_293K, 293.6), (synth.KT_1200K, 1200)):
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(kT=kT)
    dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
    for _ in range(2):
        ctx.stats(reset=True)
        el = dn.elastic(Ein)
        st = ctx.stats(reset=True)
    print("  %%6.1f K kernel %%8.2f ms  items %%d  launches %%d  sha %%s" %% (T, st["kernel_ms"], st["freegas_items"], st["launches"], hashlib.sha256(el.tobytes()).hexdigest()[:12]))
    dn.clear()
''' % ROOT
for spec in sys.argv[1:]:
    a, b, c = spec.split()
    env = dict(os.environ, NDPPGPU_FG_SPLIT=a, NDPPGPU_FG_SPLIT_LATE=b, NDPPGPU_FG_LATE_ITEMS=c)
    print("split %s, late split %s below %s items" % (a, b, c), flush=True)
    subprocess.run([sys.executable, "-c", CHILD], env=env)
