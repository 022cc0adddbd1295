import sys, time
import numpy as np
sys.path.insert(0, ".")
from ndpp_b200 import scatt, synth
ctx = scatt.default_context()
for n in (25, 50, 100, 200, 400, 1000):
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=n)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    dn.elastic(Ein[:10])
    best = 1e9
    for rep in range(2):
        ctx.stats(reset=True); dn.elastic(Ein); best = min(best, ctx.stats()["kernel_ms"])
    print(n, f"{best:.1f} ms", f"{best/n:.3f} ms/E_in")
    dn.clear()
# single E_in columns: the cost of one column alone = an upper bound of the longest task chain
nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=1000)
dn = scatt.DeviceNuclide(nuc, e_bins, params)
for k in (0, 300, 600, 900, 999):
    ctx.stats(reset=True); dn.elastic(Ein[k:k+1]); print("single E_in", k, f"{Ein[k]:.3e}", f"{ctx.stats()['kernel_ms']:.1f} ms")
