"""ORACLE loader (test infrastructure only).

ctypes front-end to oracle/libndpp_oracle.so, the plain-C restatement of the reference algorithm.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module; nothing under ndpp_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libndpp_oracle.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)


class RefParams(C.Structure):
    _fields_ = [("scatt_type", C.c_int), ("order", C.c_int), ("mu_bins", C.c_int), ("nuscatter", C.c_int),
                ("ne_per_grp", C.c_int), ("adaptive_mu_its", C.c_int), ("adaptive_eout_its", C.c_int),
                ("reserved", C.c_int), ("sab_threshold", C.c_double), ("brent_mu_thresh", C.c_double),
                ("adaptive_mu_tol", C.c_double), ("adaptive_eout_tol", C.c_double)]


def build(force: bool = False) -> str:
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in srcs):
        subprocess.check_call(["make", "-s", "-C", _HERE])
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.ref_calc_pn.restype = C.c_double
        L.ref_calc_pn.argtypes = [C.c_int, C.c_double]
        L.ref_tolab.restype = C.c_double
        L.ref_tolab.argtypes = [C.c_double, C.c_double]
        L.ref_interpolate_tab1.restype = C.c_double
        L.ref_interpolate_tab1.argtypes = [c_dp, C.c_double]
        L.ref_calc_sab.restype = C.c_double
        L.ref_calc_sab.argtypes = [C.c_double] * 6
        L.ref_calc_fgk.restype = C.c_double
        L.ref_calc_fgk.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, c_dp, c_dp, C.c_int]
        L.ref_error_message.restype = C.c_char_p
        L.ref_nuclide_create.restype = C.c_void_p
        L.ref_nuclide_create.argtypes = [C.c_double, C.c_double, C.c_double, C.c_int, c_dp, c_dp, c_dp, C.c_int,
                                         C.POINTER(RefParams)]
        L.ref_nuclide_add_reaction.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int, C.c_int,
                                               C.c_int, C.c_int, C.c_int, c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int,
                                               c_dp, c_ip, c_ip, C.c_int, c_dp, C.c_int, c_dp, C.c_int]
        for name in ("ref_nuclide_n_slots", "ref_nuclide_convert_distro", "ref_nuclide_free"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.ref_nuclide_free.restype = None
        L.ref_nuclide_slot_info.argtypes = [C.c_void_p, C.c_int, c_ip]
        L.ref_nuclide_slot_row_np.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.ref_nuclide_slot_egrid.argtypes = [C.c_void_p, C.c_int, c_dp]
        L.ref_nuclide_create_ein_grid.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_ip, c_dp, c_ip, C.c_int]
        L.ref_sab_egrid.argtypes = [c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_int, C.c_int,
                                    c_dp, C.c_int]
        L.ref_nuclide_get_table.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp, c_dp, c_dp, c_dp, c_ip]
        L.ref_nuclide_set_table.argtypes = [C.c_void_p, C.c_int, C.c_int, c_dp]
        L.ref_nuclide_interp_distro.argtypes = [C.c_void_p, C.c_int, C.c_double, c_dp]
        L.ref_nuclide_elastic.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, C.c_int]
        L.ref_nuclide_inelastic.argtypes = [C.c_void_p, c_dp, C.c_int, c_dp, c_dp, C.c_int]
        L.ref_sab_create.restype = C.c_void_p
        L.ref_sab_create.argtypes = [C.c_double] * 4 + [C.c_int] * 4 + [c_dp] * 4 + [c_ip] + [c_dp] * 3 + \
                                    [C.c_int] * 3 + [c_dp] * 3
        L.ref_sab_calc.argtypes = [C.c_void_p, c_dp, C.c_int, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp, C.c_int]
        L.ref_sab_calc_tabular.argtypes = L.ref_sab_calc.argtypes
        L.ref_sab_free.argtypes = [C.c_void_p]
        L.ref_sab_free.restype = None
        L.ref_apply_tol_scatt.argtypes = [c_dp, C.c_int, C.c_int, C.c_int, C.c_double]
        L.ref_apply_tol_scatt.restype = None
        L.ref_thin_grid.argtypes = [c_dp, c_dp, c_dp, C.c_int, C.c_int, c_dp, C.c_int, C.c_double, c_ip, c_dp, c_dp, c_dp]
        L.ref_calc_chi.argtypes = [C.c_int, c_dp, c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, c_dp, C.c_int, C.c_void_p,
                                   c_dp, c_dp, C.c_int, c_dp, C.c_int, c_dp, c_dp, c_dp]
        L.ref_freegas_counters.argtypes = [C.POINTER(C.c_longlong), C.POINTER(C.c_longlong), C.c_int]
        L.ref_freegas_counters.restype = None
        _lib = L
    return _lib


def dp(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(c_dp)


def ip(a):
    if a is None:
        return None
    assert a.dtype == np.int32 and a.flags.c_contiguous
    return a.ctypes.data_as(c_ip)


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())


def make_params(p) -> RefParams:
    return RefParams(int(p.scatt_type), int(p.order), int(p.mu_bins), int(bool(p.nuscatter)), int(p.ne_per_grp),
                     int(p.adaptive_mu_its), int(p.adaptive_eout_its), 0, float(p.sab_threshold),
                     float(p.brent_mu_thresh), float(p.adaptive_mu_tol), float(p.adaptive_eout_tol))


def check_errors():
    L = lib()
    n = L.ref_error_count()
    if n:
        msg = L.ref_error_message().decode()
        L.ref_error_clear()
        raise RuntimeError(f"oracle fatal_error x{n}: {msg}")


# ---- leaf wrappers -------------------------------------------------------------------------------
def calc_pn(n, x):
    return lib().ref_calc_pn(int(n), float(x))


def calc_int_pn_tablelin(n, xlow, xhigh, flow, fhigh):
    out = np.zeros(n)
    lib().ref_calc_int_pn_tablelin(C.c_int(n), C.c_double(xlow), C.c_double(xhigh), C.c_double(flow),
                                   C.c_double(fhigh), dp(out))
    return out


def mu_grid(M):
    mu = -1.0 + np.arange(M, dtype=np.float64) * (2.0 / float(M - 1))
    mu[-1] = 1.0
    return mu


def convert_file4(iE, mu, adist):
    mu = f64(mu)
    out = np.zeros(len(mu))
    lib().ref_convert_file4(C.c_int(iE), dp(mu), C.c_int(len(mu)), dp(f64(adist.energy)), ip(i32(adist.type)),
                            ip(i32(adist.location)), dp(f64(adist.data)), dp(out))
    return out


def convert_file6(iE, mu, law, data, NP, init=0.0):
    mu = f64(mu)
    M = len(mu)
    distro = np.full(M * NP, init)
    Eouts, pdf, cdf = np.zeros(NP), np.zeros(NP), np.zeros(NP)
    INTT, NPo = C.c_int(-1), C.c_int(0)
    rc = lib().ref_convert_file6(C.c_int(iE), dp(mu), C.c_int(M), C.c_int(law), dp(f64(data)), C.byref(INTT),
                                 C.byref(NPo), dp(Eouts), dp(pdf), dp(cdf), dp(distro))
    return rc, INTT.value, Eouts, pdf, cdf, distro.reshape(NP, M).T.copy()


def integrate_file4_cm_leg(fw, Ein, awr, Q, E_bins, w, order):
    E_bins, w, fw = f64(E_bins), f64(w), f64(fw)
    out = np.zeros((len(E_bins) - 1, order))
    lib().ref_integrate_file4_cm_leg(dp(fw), C.c_double(Ein), C.c_double(awr), C.c_double(Q), dp(E_bins),
                                     C.c_int(len(E_bins)), dp(w), C.c_int(len(w)), C.c_int(order), dp(out))
    return out  # [g][l]


def integrate_file6_lab_leg(fEmu, mu, Eout, INTT, pdf, E_bins, order):
    """fEmu: array (M, NEout) as in the Fortran; returns [g][l]."""
    mu, Eout, pdf, E_bins = f64(mu), f64(Eout), f64(pdf), f64(E_bins)
    fE = np.ascontiguousarray(np.asarray(fEmu, dtype=np.float64).T)  # [NEout][M]
    out = np.zeros((len(E_bins) - 1, order))
    lib().ref_integrate_file6_lab_leg(dp(fE.ravel()), dp(mu), C.c_int(len(mu)), dp(Eout), C.c_int(len(Eout)),
                                      C.c_int(INTT), dp(pdf), dp(E_bins), C.c_int(len(E_bins)), C.c_int(order),
                                      dp(out))
    return out


def integrate_file6_cm_leg(fEmu, mu, Ein, awr, Eout, INTT, pdf, E_bins, order, ne_per_grp=20):
    mu, Eout, pdf, E_bins = f64(mu), f64(Eout), f64(pdf), f64(E_bins)
    fE = np.ascontiguousarray(np.asarray(fEmu, dtype=np.float64).T)
    out = np.zeros((len(E_bins) - 1, order))
    lib().ref_integrate_file6_cm_leg(dp(fE.ravel()), dp(mu), C.c_int(len(mu)), C.c_double(Ein), C.c_double(awr),
                                     dp(Eout), C.c_int(len(Eout)), C.c_int(INTT), dp(pdf), dp(E_bins),
                                     C.c_int(len(E_bins)), C.c_int(order), C.c_int(ne_per_grp), dp(out))
    return out


def integrate_freegas_leg(Ein, A, kT, fEmu, mu, E_bins, order, params):
    fEmu, mu, E_bins = f64(fEmu), f64(mu), f64(E_bins)
    out = np.zeros((len(E_bins) - 1, order))
    rp = make_params(params)
    lib().ref_integrate_freegas_leg(C.c_double(Ein), C.c_double(A), C.c_double(kT), dp(fEmu), dp(mu),
                                    C.c_int(len(mu)), dp(E_bins), C.c_int(len(E_bins)), C.c_int(order), C.byref(rp),
                                    dp(out))
    return out


def freegas_counters(reset=False):
    a, b = C.c_longlong(0), C.c_longlong(0)
    lib().ref_freegas_counters(C.byref(a), C.byref(b), int(reset))
    return a.value, b.value


# ---- nuclide level ---------------------------------------------------------------------------------
class RefNuclide:
    """Oracle-side ScattData set for one nuclide: mirrors what calc_scatt builds
    (src/scatt.F90:84-126) from an ndpp_b200.ace.Nuclide."""

    def __init__(self, nuc, e_bins, params):
        from ndpp_b200.ace import iter_slots
        L = lib()
        self.L = L
        self.params = params
        self.e_bins = f64(e_bins)
        self.G = len(self.e_bins) - 1
        rp = make_params(params)
        en, el = f64(nuc.energy), f64(nuc.elastic)
        self.h = L.ref_nuclide_create(nuc.awr, nuc.kT, nuc.freegas_cutoff, len(en), dp(en), dp(el), dp(self.e_bins),
                                      len(self.e_bins), C.byref(rp))
        for idx, rxn, ed in iter_slots(nuc):
            yt = f64(rxn.multiplicity_E.flatten()) if rxn.multiplicity_E is not None else None
            sig = f64(rxn.sigma)
            ad = rxn.adist
            if ad is not None:
                ae, at, al, adata = f64(ad.energy), i32(ad.type), i32(ad.location), f64(ad.data)
            else:
                ae = at = al = adata = None
            pv = f64(ed.p_valid.flatten()) if (ed is not None and ed.p_valid is not None) else None
            edata = f64(ed.data) if ed is not None else None
            L.ref_nuclide_add_reaction(self.h, idx, rxn.MT, rxn.Q_value, rxn.threshold, int(rxn.scatter_in_cm),
                                       int(ad is not None), int(ed is not None), ed.law if ed is not None else 0,
                                       rxn.multiplicity, dp(yt), 0 if yt is None else len(yt), dp(sig), len(sig),
                                       dp(pv), 0 if pv is None else len(pv), dp(ae), ip(at), ip(al),
                                       0 if ae is None else len(ae), dp(adata), 0 if adata is None else len(adata),
                                       dp(edata), 0 if edata is None else len(edata))
        self.n_slots = L.ref_nuclide_n_slots(self.h)
        self.order_L = 0
        for s in range(self.n_slots):
            info = self.slot_info(s)
            if info["is_init"]:
                self.order_L = info["order"]
        check_errors()

    def slot_info(self, s):
        info = (C.c_int * 8)()
        self.L.ref_nuclide_slot_info(self.h, s, info)
        keys = ("is_init", "NE", "law", "has_adist", "has_edist", "order", "groups", "MT")
        return dict(zip(keys, list(info)))

    def convert_distro(self):
        self.L.ref_nuclide_convert_distro(self.h)
        check_errors()

    def get_table(self, s, iE):
        """Row iE (1-based) of slot s: (distro (M, NP), Eouts, pdf, cdf, INTT)."""
        NP = self.L.ref_nuclide_slot_row_np(self.h, s, iE)
        M = self.params.mu_bins
        d = np.zeros(M * NP)
        Eo, pdf, cdf = np.zeros(NP), np.zeros(NP), np.zeros(NP)
        INTT = C.c_int(0)
        self.L.ref_nuclide_get_table(self.h, s, iE, dp(d), dp(Eo), dp(pdf), dp(cdf), C.byref(INTT))
        return d.reshape(NP, M).T.copy(), Eo, pdf, cdf, INTT.value

    def create_ein_grid(self, extend_pts=50, inel_extend_pts=30):
        """create_Ein_grid (src/scatt.F90:166-236) with calc_scatt's inel_thresh / cutoff: (Ein_el, Ein_inel or None)."""
        cap = 1 << 16
        while True:
            el, inel = np.zeros(cap), np.zeros(cap)
            n_el, n_inel = C.c_int(0), C.c_int(0)
            rc = self.L.ref_nuclide_create_ein_grid(self.h, extend_pts, inel_extend_pts, dp(el), C.byref(n_el), dp(inel),
                                                    C.byref(n_inel), cap)
            if rc == 0:
                return el[:n_el.value].copy(), (inel[:n_inel.value].copy() if n_inel.value else None)
            cap = 2 * max(n_el.value, n_inel.value)

    def slot_egrid(self, s):
        ne = self.slot_info(s)["NE"]
        out = np.zeros(ne)
        self.L.ref_nuclide_slot_egrid(self.h, s, dp(out))
        return out

    def interp_distro(self, s, Ein):
        out = np.zeros((self.G, self.order_L))
        rc = self.L.ref_nuclide_interp_distro(self.h, s, float(Ein), dp(out))
        assert rc == 0
        check_errors()
        return out

    def elastic(self, Ein, n_threads=1):
        Ein = f64(Ein)
        out = np.zeros((len(Ein), self.G, self.order_L))
        self.L.ref_nuclide_elastic(self.h, dp(Ein), len(Ein), dp(out), n_threads)
        check_errors()
        return out

    def inelastic(self, Ein, n_threads=1):
        Ein = f64(Ein)
        out = np.zeros((len(Ein), self.G, self.order_L))
        nu = np.zeros_like(out) if self.params.nuscatter else None
        self.L.ref_nuclide_inelastic(self.h, dp(Ein), len(Ein), dp(out), dp(nu), n_threads)
        check_errors()
        return out, nu

    def close(self):
        if self.h:
            self.L.ref_nuclide_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sab_egrid(sab, e_bins, sab_epts_per_bin=10, extend_pts=50):
    """sab_egrid (src/sab.F90:460-568)."""
    L = lib()
    e_bins = f64(e_bins)
    ein = f64(sab.inelastic_e_in)
    if int(sab.secondary_mode) == 2:                   # continuous secondary energies: no crossing points (:497)
        n_eo, eo = 0, None
    else:
        n_eo = np.asarray(sab.inelastic_e_out).reshape(len(ein), -1).shape[1]
        eo = f64(sab.inelastic_e_out)                  # [n_e_in][n_e_out], C order = the Fortran (j, i) layout
    eel = f64(sab.elastic_e_in) if sab.elastic_e_in is not None else None
    cap = 1 << 16
    while True:
        out = np.zeros(cap)
        n = L.ref_sab_egrid(dp(ein), len(ein), dp(eel), 0 if eel is None else len(eel), dp(eo), n_eo,
                            int(sab.secondary_mode), dp(e_bins), len(e_bins), sab_epts_per_bin, extend_pts, dp(out), cap)
        if n >= 0:
            return out[:n].copy()
        cap = -n


def sab_calc(sab, e_bins, order, Ein, parts=False, tabular=False):
    """calc_scattsab (src/scatt.F90:543-596) on an ndpp_b200.ace.SAlphaBeta; returns [iE][g][l]
    (tabular: `order` cosine bins, project-defined semantics, see sab_ref.c)."""
    from ndpp_b200.ace import SAB_SECONDARY_CONT
    L = lib()
    e_bins, Ein = f64(e_bins), f64(Ein)
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
    h = L.ref_sab_create(sab.awr, sab.kT, sab.threshold_inelastic, sab.threshold_elastic, sab.n_inelastic_e_in, neo,
                         sab.n_inelastic_mu, sab.secondary_mode, dp(ei), dp(sg), dp(eo), dp(mu), ip(cn), dp(ce),
                         dp(cp), dp(cm), sab.elastic_mode, sab.n_elastic_e_in, sab.n_elastic_mu, dp(ee), dp(eP),
                         dp(em))
    G = len(e_bins) - 1
    out = np.zeros((len(Ein), G, order if tabular else order + 1))
    el, inel = np.zeros_like(out), np.zeros_like(out)
    (L.ref_sab_calc_tabular if tabular else L.ref_sab_calc)(h, dp(e_bins), len(e_bins), order, dp(Ein), len(Ein), dp(out), dp(el), dp(inel), 1)
    L.ref_sab_free(h)
    check_errors()
    return (out, el, inel) if parts else out


def apply_tol_scatt(data, tol):
    """apply_tol_scatt (src/scatt.F90:786-818) on [iE][g][l]; returns a new array."""
    d = np.array(data, dtype=np.float64, order="C", copy=True)
    NE, G, Lm = d.shape
    lib().ref_apply_tol_scatt(dp(d), Lm, G, NE, float(tol))
    return d


def thin_grid(x, y, tokeep, tol, y2=None):
    """thin_grid (src/thin.F90): returns (kept indices, compression, maxerr as the reference computes it,
    plain max |interpolated - y|)."""
    x, y = f64(x), f64(y)
    NE = len(x)
    GL = y.size // max(NE, 1)
    y2c = f64(y2) if y2 is not None else None
    tk = f64(tokeep)
    keep = np.zeros(max(NE, 1), np.int32)
    c, m, a = C.c_double(0), C.c_double(0), C.c_double(0)
    n = lib().ref_thin_grid(dp(x), dp(y), dp(y2c), NE, GL, dp(tk), len(tk), float(tol), ip(keep), C.byref(c), C.byref(m),
                            C.byref(a))
    return keep[:n].copy(), c.value, m.value, a.value


def calc_chi(nuc, E_bins, E_grid=None):
    """calc_chi (src/chi.F90:21-163) through the C restatement; the slot list, pool and merged grid are the
    host mirror's (ndpp_b200.chi: data marshalling only, no arithmetic on the moments)."""
    from ndpp_b200 import chi as hostchi
    slots, pool = hostchi.chi_data(nuc)
    E_grid = hostchi.chi_grid(slots) if E_grid is None else f64(E_grid)
    E_bins = f64(E_bins)
    G, NE, n_prec = len(E_bins) - 1, len(E_grid), nuc.n_precursor
    energy, fis = f64(nuc.energy), hostchi.fission_xs(nuc)
    nu_t = f64(nuc.nu_t_data)
    nu_d = f64(nuc.nu_d_data) if nuc.nu_d_data is not None else np.zeros(1)
    prec = f64(nuc.nu_d_precursor_data) if nuc.nu_d_precursor_data is not None else np.zeros(1)
    chi_t, chi_p, chi_d = np.empty((NE, G)), np.empty((NE, G)), np.empty((max(n_prec, 1), NE, G))
    sc = hostchi.slots_c(slots)
    rc = lib().ref_calc_chi(len(energy), dp(energy), dp(fis), int(nuc.nu_t_type), dp(nu_t), int(nuc.nu_d_type), dp(nu_d),
                            n_prec, dp(prec), len(slots), C.cast(sc, C.c_void_p), dp(pool), dp(E_bins), len(E_bins),
                            dp(E_grid), NE, dp(chi_t), dp(chi_p), dp(chi_d))
    if rc:
        check_errors()
    return E_grid, chi_t, chi_p, chi_d[:n_prec]
