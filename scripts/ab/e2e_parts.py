"""Splits an end-to-end pass of the bench workload (C2) into its phases: DeviceNuclide(...) (ctypes), convert_distro, calc
(ndppgpu_calc_scatt with page-locked host arrays) and clear.  Run on the GPU box: python scripts/ab/e2e_parts.py"""
import sys, time, numpy as np, torch
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
from ndpp_b200 import scatt
from ndpp_b200.capi import Context
nuc, e_bins, params, Eel, Einel = bench.make_workload(20000)
ctx = Context(0)
GL = (len(e_bins)-1)*(params.order+1)
h_el = torch.empty((len(Eel), GL), dtype=torch.float64).pin_memory().numpy()
h_in = torch.empty((len(Einel), GL), dtype=torch.float64).pin_memory().numpy()
for rep in range(4):
    ctx.stats(reset=True)
    t0=time.perf_counter(); dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx, convert=False) if 'convert' in scatt.DeviceNuclide.__init__.__code__.co_varnames else scatt.DeviceNuclide(nuc, e_bins, params, ctx)
    t1=time.perf_counter(); 
    if hasattr(dn,'convert_distro') and 'convert' in scatt.DeviceNuclide.__init__.__code__.co_varnames: dn.convert_distro()
    torch.cuda.synchronize(); t2=time.perf_counter()
    dn.calc(Eel, Einel, False, el_out=h_el, inel_out=h_in); t3=time.perf_counter()
    dn.clear(); t4=time.perf_counter()
    st=ctx.stats()
    print('create %.2f convert %.2f calc %.2f clear %.2f ms | kernel %.2f host_alloc %.2f host_call %.2f h2d %.1f MB' % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3, st['kernel_ms'], st.get('host_alloc_ms',0), st.get('host_call_ms',0), st['h2d_bytes']/1e6))
