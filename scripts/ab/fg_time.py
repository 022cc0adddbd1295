#!/usr/bin/env python
"""Times the free-gas path (C3, 293.6 K and 1200 K, 1000 E_in; second pass) on every library variant in $VARIANTS
(scripts/ab/libs/NAME.so, built by ab_build.py) and prints kernel ms, work items and a hash of the moments."""
import hashlib
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

CHILD = r'''
import sys, hashlib, numpy as np
sys.path.insert(0, %r)
from ndpp_b200 import scatt, synth
ctx = scatt.default_context()
for kT, T in ((synth.KT_293K, 293.6), (synth.KT_1200K, 1200)):
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(kT=kT)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else len(Ein)
    Ein = Ein[:: max(1, len(Ein) // n)]
    dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
    for _ in range(2):
        ctx.stats(reset=True)
        el = dn.elastic(Ein)
        st = ctx.stats(reset=True)
    print("  %%6.1f K  %%4d E_in  kernel %%8.2f ms  items %%d  launches %%d  sha %%s" %% (T, len(Ein), st["kernel_ms"], st["freegas_items"], st["launches"], hashlib.sha256(el.tobytes()).hexdigest()[:12]))
    dn.clear()
''' % ROOT

so = os.path.join(ROOT, "ndpp_b200", "csrc", "libndppgpu.so")
shutil.copy(so, "/tmp/libndppgpu.default.so")
try:
    for v in os.environ.get("VARIANTS", "").split():
        shutil.copy(os.path.join(ROOT, "scripts", "ab", "libs", v + ".so"), so)
        print(v, flush=True)
        subprocess.run([sys.executable, "-c", CHILD] + sys.argv[1:])
finally:
    shutil.copy("/tmp/libndppgpu.default.so", so)
