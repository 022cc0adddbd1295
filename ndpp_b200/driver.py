"""Host-side mirror of the per-nuclide body of the reference's driver, src/ndpp.F90:560-702
(`preprocess_ndpp`): E_in grids -> calc_scatt -> apply_tol_scatt -> thin_grid -> group indices ->
library file.  Everything numerical runs in libndppgpu.so; the tolerance and the thinning run on the
device (ndppgpu_*_thinned), so only the columns that end up in the library cross PCIe.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import chi, output
from .ace import SCATT_TYPE_LEGENDRE, Nuclide, Params
from .capi import Context, check, dp, f64
from .scatt import DeviceNuclide

PRINT_TOL_DEFAULT = 1.0e-8   # src/constants.F90 (print_tol default)


@dataclass
class NuclideResult:
    Ein_el: np.ndarray
    el_mat: np.ndarray
    Ein_inel: Optional[np.ndarray]
    inel_mat: Optional[np.ndarray]
    nuinel_mat: Optional[np.ndarray]
    thin_compr_el: float = 0.0
    thin_err_el: float = 0.0
    thin_compr_inel: float = 0.0
    thin_err_inel: float = 0.0
    chi: Optional[tuple] = None   # (E_grid, chi_total, chi_prompt, chi_delay) when integrate_chi and fissionable


def _thinned(dn: DeviceNuclide, inelastic: bool, Ein, print_tol, thin_tol, tokeep, nuscatter):
    x = np.array(Ein, dtype=np.float64, order="C", copy=True)
    NE = len(x)
    mat = np.empty((NE, dn.G, dn.L))
    nu = np.empty_like(mat) if (inelastic and nuscatter) else None
    tk = f64(tokeep)
    n, comp, err = C.c_int(0), C.c_double(0.0), C.c_double(0.0)
    if inelastic:
        check(dn.lib.ndppgpu_inelastic_thinned(dn.h, dp(x), NE, float(print_tol), float(thin_tol), dp(tk), len(tk), dp(mat),
                                               dp(nu), C.byref(n), C.byref(comp), C.byref(err)), dn.ctx.h)
    else:
        check(dn.lib.ndppgpu_elastic_thinned(dn.h, dp(x), NE, float(print_tol), float(thin_tol), dp(tk), len(tk), dp(mat),
                                             C.byref(n), C.byref(comp), C.byref(err)), dn.ctx.h)
    k = n.value
    return x[:k], mat[:k], (nu[:k] if nu is not None else None), comp.value, err.value


def preprocess_nuclide(nuc: Nuclide, energy_bins, params: Params, print_tol: float = PRINT_TOL_DEFAULT,
                       thin_tol: float = 0.0, Ein_el=None, Ein_inel=None, ctx: Optional[Context] = None,
                       library_file: Optional[str] = None, lib_format: str = output.BINARY,
                       integrate_chi: bool = False) -> NuclideResult:
    """One nuclide through src/ndpp.F90:560-702.  `thin_tol` is the fraction the reference derives from the
    user's percentage (`0.01 * thinning_tol`, :337); E_in grids default to create_Ein_grid (src/scatt.F90:166), built on
    the device after convert_distro as in calc_scatt (:107-139)."""
    if params.scatt_type != SCATT_TYPE_LEGENDRE:
        raise ValueError("tabular scattering of ACE nuclides is NOT YET IMPLEMENTED in the reference")
    dn = DeviceNuclide(nuc, energy_bins, params, ctx)
    try:
        if Ein_el is None:
            Ein_el, Ein_inel, _ = dn.create_ein_grid()
        xe, el, _, ce, ee = _thinned(dn, False, Ein_el, print_tol, thin_tol, energy_bins, False)
        xi = inel = nu = None
        ci = ei = 0.0
        if Ein_inel is not None and len(Ein_inel) > 0:
            xi, inel, nu, ci, ei = _thinned(dn, True, Ein_inel, print_tol, thin_tol, energy_bins, params.nuscatter)
    finally:
        dn.clear()
    res = NuclideResult(xe, el, xi, inel, nu, ce, ee, ci, ei)
    do_chi = bool(integrate_chi) and nuc.fissionable               # src/ndpp.F90:705-712
    if do_chi:
        res.chi = chi.calc_chi(nuc, energy_bins, ctx)
    if library_file is not None:
        with output.LibraryWriter(library_file, nuc.name, nuc.kT, energy_bins, params.scatt_type, params.order,
                                  params.nuscatter, params.mu_bins, thin_tol, lib_format, chi_present=do_chi) as w:
            w.print_scatt(xe, el, xi, inel, nu)
            if do_chi:
                w.print_chi(*res.chi)
    return res
