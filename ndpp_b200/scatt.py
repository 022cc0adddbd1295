"""Host-side mirror of the reference's scattering orchestration, src/scatt.F90.

`calc_scatt` and `calc_scattsab` keep the reference's names, argument meaning and error behaviour
(src/scatt.F90:33-46, 543-552); their bodies hand the parsed nuclide to libndppgpu.so through the
C-ABI (include/ndppgpu.h) exactly as the Fortran shim of INTEGRATION.md does.  All numerical work
happens in the CUDA library; a missing library or GPU raises NdppGpuError.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import capi
from .ace import SAB_SECONDARY_CONT, SCATT_TYPE_LEGENDRE, Nuclide, Params, SAlphaBeta, iter_slots
from .capi import Context, NdppGpuError, check, dp, f64, i32, ip

_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(-1)
    return _default_ctx


def reaction_args(nuc: Nuclide):
    """The argument lists of ndppgpu_nuclide_add_reaction (after the handle) for every ScattData slot of a nuclide, in the
    order calc_scatt fills rxn_data(:) (src/scatt.F90:88-105)."""
    for idx, rxn, ed in iter_slots(nuc):
        yt = f64(rxn.multiplicity_E.flatten()) if rxn.multiplicity_E is not None else None
        sig = f64(rxn.sigma)
        ad = rxn.adist
        ae = at = al = adata = None
        if ad is not None:
            ae, at, al, adata = f64(ad.energy), i32(ad.type), i32(ad.location), f64(ad.data)
        pv = f64(ed.p_valid.flatten()) if (ed is not None and ed.p_valid is not None) else None
        edata = f64(ed.data) if ed is not None else None
        yield (idx, rxn.MT, rxn.Q_value, rxn.threshold, int(rxn.scatter_in_cm), int(ad is not None),
               int(ed is not None), ed.law if ed is not None else 0, rxn.multiplicity, dp(yt),
               0 if yt is None else len(yt), dp(sig), len(sig), dp(pv), 0 if pv is None else len(pv), dp(ae), ip(at),
               ip(al), 0 if ae is None else len(ae), dp(adata), 0 if adata is None else len(adata), dp(edata),
               0 if edata is None else len(edata))


class DeviceNuclide:
    """Device-resident ScattData set of one nuclide: what calc_scatt builds in rxn_data(:)
    (src/scatt.F90:84-126) -- scatt_init for every (reaction, energy distribution) slot followed by
    convert_distro."""

    def __init__(self, nuc: Nuclide, energy_bins, params: Params, ctx: Optional[Context] = None, convert=True):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        self.params = params
        self.e_bins = f64(energy_bins)
        self.G = len(self.e_bins) - 1
        self.L = params.order + 1 if params.scatt_type == SCATT_TYPE_LEGENDRE else params.order
        self.h = C.c_void_p()
        en, el = f64(nuc.energy), f64(nuc.elastic)
        pc = capi.make_params(params)
        check(self.lib.ndppgpu_nuclide_create(self.ctx.h, nuc.awr, nuc.kT, nuc.freegas_cutoff, len(en), dp(en), dp(el),
                                              dp(self.e_bins), len(self.e_bins), C.byref(pc), C.byref(self.h)),
              self.ctx.h)
        for args in reaction_args(nuc):
            check(self.lib.ndppgpu_nuclide_add_reaction(self.h, *args), self.ctx.h)
        self.n_slots = self.lib.ndppgpu_nuclide_n_slots(self.h)
        if convert:
            self.convert_distro()

    # -- ScattData % convert_distro ------------------------------------------------------------
    def convert_distro(self):
        check(self.lib.ndppgpu_convert_distro(self.h), self.ctx.h)

    def slot_info(self, s):
        info = (C.c_int * 8)()
        check(self.lib.ndppgpu_nuclide_slot_info(self.h, s, info), self.ctx.h)
        keys = ("is_init", "NE", "law", "has_adist", "has_edist", "order", "groups", "MT")
        return dict(zip(keys, list(info)))

    def get_table(self, s, iE):
        NP = self.lib.ndppgpu_nuclide_slot_row_np(self.h, s, iE)
        if NP < 0:
            raise NdppGpuError("get_table: bad slot / row")
        M = self.params.mu_bins
        d = np.zeros(M * NP)
        Eo, pdf, cdf = np.zeros(NP), np.zeros(NP), np.zeros(NP)
        INTT = C.c_int(0)
        check(self.lib.ndppgpu_nuclide_get_table(self.h, s, iE, dp(d), dp(Eo), dp(pdf), dp(cdf), C.byref(INTT)),
              self.ctx.h)
        return d.reshape(NP, M).T.copy(), Eo, pdf, cdf, INTT.value

    def set_table(self, s, iE, distro):
        """Overwrite row iE (1-based) of slot s with distro (M, NP) as get_table returns it."""
        d = np.ascontiguousarray(np.asarray(distro, dtype=np.float64).T)
        check(self.lib.ndppgpu_nuclide_set_table(self.h, s, iE, dp(d.ravel())), self.ctx.h)

    def interp_distro(self, s, Ein) -> np.ndarray:
        """mySD % interp_distro (src/scattdata_header.F90:391) for slot s at every E_in: [NE][G][L]."""
        Ein = f64(Ein)
        out = np.empty((len(Ein), self.G, self.L))
        check(self.lib.ndppgpu_interp_distro(self.h, s, dp(Ein), len(Ein), dp(out)), self.ctx.h)
        return out

    # -- calc_elastic_grid / calc_inelastic_grid, host buffers -----------------------------------
    def _out(self, out, n):
        """Result buffer [n][G][L]; a caller-owned (e.g. page-locked) array may be passed in, as the
        Fortran caller passes its own allocatable (src/scatt.F90:628,711)."""
        if out is None:
            return np.empty((n, self.G, self.L))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == n * self.G * self.L
        return out

    def elastic(self, Ein, out=None) -> np.ndarray:
        Ein = f64(Ein)
        out = self._out(out, len(Ein))
        check(self.lib.ndppgpu_elastic(self.h, dp(Ein), len(Ein), dp(out)), self.ctx.h)
        return out

    def inelastic(self, Ein, nuscatt=None, out=None, nu_out=None):
        Ein = f64(Ein)
        nuscatt = self.params.nuscatter if nuscatt is None else nuscatt
        out = self._out(out, len(Ein))
        nu = self._out(nu_out, len(Ein)) if nuscatt else None
        check(self.lib.ndppgpu_inelastic(self.h, dp(Ein), len(Ein), dp(out), dp(nu)), self.ctx.h)
        return out, nu

    def calc(self, Ein_el, Ein_inel=None, nuscatt=None, el_out=None, inel_out=None, nu_out=None):
        """elastic + inelastic in one call (ndppgpu_calc_scatt): the elastic matrices are copied to the host while the
        inelastic kernels run.  Returns (el, inel, nu); inel / nu are None without an inelastic grid."""
        Eel = f64(Ein_el)
        nuscatt = self.params.nuscatter if nuscatt is None else nuscatt
        el = self._out(el_out, len(Eel))
        Ein = inel = nu = None
        if Ein_inel is not None and len(Ein_inel) > 0:
            Ein = f64(Ein_inel)
            inel = self._out(inel_out, len(Ein))
            nu = self._out(nu_out, len(Ein)) if nuscatt else None
        check(self.lib.ndppgpu_calc_scatt(self.h, dp(Eel), len(Eel), dp(el), dp(Ein), 0 if Ein is None else len(Ein),
                                          dp(inel), dp(nu)), self.ctx.h)
        return el, inel, nu

    # -- device-resident variants (torch CUDA tensors of dtype float64) --------------------------
    def elastic_dev(self, d_Ein, d_out):
        check(self.lib.ndppgpu_elastic_dev(self.h, d_Ein.data_ptr(), int(d_Ein.numel()), d_out.data_ptr()), self.ctx.h)

    def inelastic_dev(self, d_Ein, d_out, d_nu=None):
        check(self.lib.ndppgpu_inelastic_dev(self.h, d_Ein.data_ptr(), int(d_Ein.numel()), d_out.data_ptr(),
                                             d_nu.data_ptr() if d_nu is not None else None), self.ctx.h)

    def create_ein_grid(self, extend_pts=50, inel_extend_pts=30, host=True):
        """create_Ein_grid (src/scatt.F90:166-236) on the device.  Returns (Ein_el, Ein_inel or None, status) as host
        arrays, or -- host=False -- ((device pointer, n_el), (device pointer, n_inel) or None, status) for the *_dev calls."""
        n_el, n_inel, st = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.lib.ndppgpu_nuclide_create_ein_grid(self.h, extend_pts, inel_extend_pts, C.byref(n_el), C.byref(n_inel),
                                                       C.byref(st)), self.ctx.h)
        out = []
        for which, n in ((0, n_el.value), (1, n_inel.value)):
            if n == 0:
                out.append(None)
                continue
            ptr = C.c_void_p()
            arr = np.empty(n) if host else None
            check(self.lib.ndppgpu_nuclide_ein_grid(self.h, which, dp(arr), C.byref(ptr)), self.ctx.h)
            out.append(arr if host else (ptr.value, n))
        return out[0], out[1], st.value

    def clear(self):
        """rxn_data(i) % clear() (src/scatt.F90:153-155)."""
        if self.h:
            self.lib.ndppgpu_nuclide_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


def calc_scatt(nuc: Nuclide, energy_bins, scatt_type: int, order: int, mu_bins: int, nuscatt: bool, Ein_el, Ein_inel,
               params: Optional[Params] = None, ctx: Optional[Context] = None):
    """calc_scatt (src/scatt.F90:33-157).  Returns (el_mat, inel_mat, nuinel_mat) as arrays
    [NE][G][L] (the Fortran mat(L, G, NE)); inel_mat / nuinel_mat are None when Ein_inel is None,
    as they stay unallocated in the reference (src/scatt.F90:146-150).

    The E_in grids are inputs: create_Ein_grid (src/scatt.F90:166-536) stays on the host side of the
    seam and hands its result in."""
    p = params or Params()
    p = Params(**{**p.__dict__, "scatt_type": scatt_type, "order": order, "mu_bins": mu_bins, "nuscatter": nuscatt})
    dn = DeviceNuclide(nuc, energy_bins, p, ctx)
    try:
        el, inel, nu = dn.calc(Ein_el, Ein_inel, nuscatt)
    finally:
        dn.clear()
    return el, inel, nu


class DeviceSab:
    """Device-resident S(a,b) table (type(SAlphaBeta), src/ace_header.F90:201-235)."""

    def __init__(self, sab: SAlphaBeta, ctx: Optional[Context] = None):
        self.ctx = ctx or default_context()
        self.lib = self.ctx.lib
        ei, sg = f64(sab.inelastic_e_in), f64(sab.inelastic_sigma)
        eo = mu = cn = ce = cp = cm = None
        neo = 0
        if sab.secondary_mode == SAB_SECONDARY_CONT:
            cn = i32([len(d.e_out) for d in sab.inelastic_data])
            ce = f64(np.concatenate([d.e_out for d in sab.inelastic_data]))
            cp = f64(np.concatenate([d.e_out_pdf for d in sab.inelastic_data]))
            cm = f64(np.concatenate([np.asarray(d.mu).ravel() for d in sab.inelastic_data]))
        else:
            eo, mu = f64(sab.inelastic_e_out), f64(sab.inelastic_mu)
            neo = sab.n_inelastic_e_out
        ee = f64(sab.elastic_e_in) if sab.elastic_e_in is not None else None
        eP = f64(sab.elastic_P) if sab.elastic_P is not None else None
        em = f64(sab.elastic_mu) if sab.elastic_mu is not None else None
        self.h = C.c_void_p()
        check(self.lib.ndppgpu_sab_create(self.ctx.h, sab.awr, sab.kT, sab.threshold_inelastic, sab.threshold_elastic,
                                          sab.n_inelastic_e_in, neo, sab.n_inelastic_mu, sab.secondary_mode, dp(ei),
                                          dp(sg), dp(eo), dp(mu), ip(cn), dp(ce), dp(cp), dp(cm), sab.elastic_mode,
                                          sab.n_elastic_e_in, sab.n_elastic_mu, dp(ee), dp(eP), dp(em),
                                          C.byref(self.h)), self.ctx.h)

    def calc(self, energy_bins, scatt_type, order, E_grid, parts=False):
        eb, Ein = f64(energy_bins), f64(E_grid)
        out = np.empty((len(Ein), len(eb) - 1, order + 1 if scatt_type == SCATT_TYPE_LEGENDRE else order))
        el = np.empty_like(out) if parts else None
        inel = np.empty_like(out) if parts else None
        check(self.lib.ndppgpu_sab(self.h, dp(eb), len(eb), scatt_type, order, dp(Ein), len(Ein), dp(out), dp(el),
                                   dp(inel)), self.ctx.h)
        return (out, el, inel) if parts else out

    def egrid(self, energy_bins, sab_epts_per_bin=10, extend_pts=50):
        """sab_egrid (src/sab.F90:460-568) on the device: (Ein, status)."""
        eb = f64(energy_bins)
        n, st = C.c_int(0), C.c_int(0)
        check(self.lib.ndppgpu_sab_egrid(self.h, dp(eb), len(eb), sab_epts_per_bin, extend_pts, C.byref(n), C.byref(st)),
              self.ctx.h)
        out = np.empty(n.value)
        check(self.lib.ndppgpu_sab_ein_grid(self.h, dp(out), None), self.ctx.h)
        return out, st.value

    def calc_dev(self, energy_bins, scatt_type, order, d_Ein, d_out):
        eb = f64(energy_bins)
        check(self.lib.ndppgpu_sab_dev(self.h, dp(eb), len(eb), scatt_type, order, d_Ein.data_ptr(),
                                       int(d_Ein.numel()), d_out.data_ptr()), self.ctx.h)

    def clear(self):
        if self.h:
            self.lib.ndppgpu_sab_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


def calc_scattsab(sab: SAlphaBeta, energy_bins, scatt_type: int, order: int, mu_bins: int, E_grid,
                  ctx: Optional[Context] = None) -> np.ndarray:
    """calc_scattsab (src/scatt.F90:543-596): scatt_mat[NE][G][order+1] (Legendre) or [NE][G][order] cosine bins
    (tabular: a TODO in the reference, :579-588; semantics in DESIGN.md).  E_grid comes from sab_egrid
    (src/sab.F90:460), which stays on the host side of the seam.  mu_bins is accepted and unused, as
    in the reference."""
    ds = DeviceSab(sab, ctx)
    try:
        return ds.calc(energy_bins, scatt_type, order, E_grid)
    finally:
        ds.clear()


# ---- the steps that follow the integrator in the reference's driver (src/ndpp.F90:611-648) --------
def apply_tol_scatt(data: np.ndarray, tol: float, ctx: Optional[Context] = None) -> np.ndarray:
    """apply_tol_scatt(data, tol) (src/scatt.F90:786-818) on [NE][G][L], in place like the Fortran."""
    ctx = ctx or default_context()
    assert data.dtype == np.float64 and data.flags.c_contiguous and data.ndim == 3
    NE, G, L = data.shape
    check(ctx.lib.ndppgpu_apply_tol(ctx.h, dp(data), NE, G, L, float(tol)), ctx.h)
    return data


def thin_grid(xout: np.ndarray, yout: np.ndarray, tokeep, tol: float, yout2: Optional[np.ndarray] = None,
              ctx: Optional[Context] = None):
    """thin_grid(xout, yout, tokeep, tol, compression, maxerr [, yout2]) (src/thin.F90:19-47).  Returns
    (xout, yout, yout2, compression, max_abs_err) with the arrays cut to the points kept (the Fortran
    re-allocates them).  max_abs_err is the plain max |interpolated - y| (include/ndppgpu.h)."""
    ctx = ctx or default_context()
    x = np.array(xout, dtype=np.float64, order="C", copy=True)
    y = np.array(yout, dtype=np.float64, order="C", copy=True)
    y2 = np.array(yout2, dtype=np.float64, order="C", copy=True) if yout2 is not None else None
    NE = len(x)
    GL = y.size // max(NE, 1)
    tk = f64(tokeep)
    n = C.c_int(0)
    comp, merr = C.c_double(0.0), C.c_double(0.0)
    check(ctx.lib.ndppgpu_thin_grid(ctx.h, dp(x), dp(y), dp(y2), NE, GL, dp(tk), len(tk), float(tol), C.byref(n),
                                    C.byref(comp), C.byref(merr)), ctx.h)
    k = n.value
    y = y.reshape(NE, -1)[:k].reshape((k,) + yout.shape[1:])
    if y2 is not None:
        y2 = y2.reshape(NE, -1)[:k].reshape((k,) + yout2.shape[1:])
    return x[:k], y, y2, comp.value, merr.value
