"""Flat binary case files exchanged with the C++ host layer (include/ndpp_host.hpp, tools/ndpp_calc_scatt.cpp).

ACE parsing stays in the reference's Fortran, so the C++ driver needs its *parsed* input in some form:
a case file is the argument list of `calc_scatt` (src/scatt.F90:33-46) or `calc_scattsab` (:543-552) written as
one stream of little-endian float64 values (integers stored as exact doubles), field after field in the
order of the reference's derived types.  The reader is `read_case` in tools/ndpp_calc_scatt.cpp.

    case   := KIND_NUCLIDE nuclide e_bins scatt_type order mu_bins nuscatt settings vec(Ein_el) vec(Ein_inel)
            | KIND_SAB     sab     e_bins scatt_type order mu_bins vec(E_grid)
    vec    := n x[n]
    tab1   := present [vec(nbt) vec(int) vec(x) vec(y)]
    result := KIND NE_el G L NE_inel has_nu el_mat[NE_el*G*L] inel_mat[...] nuinel_mat[...]
"""
from __future__ import annotations

import numpy as np

from .ace import SAB_SECONDARY_CONT, Nuclide, Params, SAlphaBeta

KIND_NUCLIDE, KIND_SAB = 1.0, 2.0


class _Out:
    def __init__(self):
        self.parts = []

    def num(self, *v):
        self.parts.append(np.asarray(v, dtype=np.float64))

    def vec(self, a):
        a = np.zeros(0) if a is None else np.asarray(a, dtype=np.float64).ravel()
        self.num(len(a))
        self.parts.append(a)

    def tab1(self, t):
        if t is None:
            self.num(0)
            return
        self.num(1)
        self.vec(t.nbt)
        self.vec(t.int)
        self.vec(t.x)
        self.vec(t.y)

    def save(self, path):
        np.concatenate(self.parts).astype("<f8").tofile(path)


def _settings(o: _Out, p: Params):
    o.num(p.ne_per_grp, p.adaptive_mu_its, p.adaptive_eout_its, p.sab_threshold, p.brent_mu_thresh,
          p.adaptive_mu_tol, p.adaptive_eout_tol)


def write_nuclide_case(path, nuc: Nuclide, energy_bins, scatt_type, order, mu_bins, nuscatt, Ein_el, Ein_inel,
                       params: Params | None = None):
    p = params or Params()
    o = _Out()
    o.num(KIND_NUCLIDE, nuc.awr, nuc.kT, nuc.freegas_cutoff)
    o.vec(nuc.energy)
    o.vec(nuc.elastic)
    o.num(len(nuc.reactions))
    for r in nuc.reactions:
        o.num(r.MT, r.Q_value, r.multiplicity, r.threshold, int(r.scatter_in_cm))
        o.vec(r.sigma)
        o.tab1(r.multiplicity_E)
        o.num(int(r.adist is not None))
        if r.adist is not None:
            o.vec(r.adist.energy)
            o.vec(r.adist.type)
            o.vec(r.adist.location)
            o.vec(r.adist.data)
        chain = []
        ed = r.edist
        while ed is not None:
            chain.append(ed)
            ed = ed.next
        o.num(len(chain))
        for ed in chain:
            o.num(ed.law)
            o.vec(ed.data)
            o.tab1(ed.p_valid)
    o.vec(energy_bins)
    o.num(scatt_type, order, mu_bins, int(bool(nuscatt)))
    _settings(o, p)
    o.vec(Ein_el)
    o.vec(Ein_inel)
    o.save(path)


def write_sab_case(path, sab: SAlphaBeta, energy_bins, scatt_type, order, mu_bins, E_grid):
    o = _Out()
    o.num(KIND_SAB, sab.awr, sab.kT, sab.threshold_inelastic, sab.threshold_elastic, sab.n_inelastic_e_in,
          sab.n_inelastic_e_out, sab.n_inelastic_mu, sab.secondary_mode)
    o.vec(sab.inelastic_e_in)
    o.vec(sab.inelastic_sigma)
    o.vec(sab.inelastic_e_out)
    o.vec(sab.inelastic_mu)
    rows = sab.inelastic_data if sab.secondary_mode == SAB_SECONDARY_CONT else []
    o.num(len(rows))
    for d in rows:
        o.vec(d.e_out)
        o.vec(d.e_out_pdf)
        o.vec(d.mu)
    o.num(sab.elastic_mode, sab.n_elastic_e_in, sab.n_elastic_mu)
    o.vec(sab.elastic_e_in)
    o.vec(sab.elastic_P)
    o.vec(sab.elastic_mu)
    o.vec(energy_bins)
    o.num(scatt_type, order, mu_bins)
    o.vec(E_grid)
    o.save(path)


def write_result(path, el, inel=None, nu=None, kind=KIND_NUCLIDE):
    """The result-file layout of tools/ndpp_calc_scatt.cpp, from [NE][G][L] arrays (used to feed its library writer)."""
    el = np.asarray(el, dtype=np.float64)
    _, G, L = el.shape
    ne_in = 0 if inel is None else len(inel)
    parts = [np.array([kind, len(el), G, L, ne_in, 0.0 if nu is None else 1.0]), el.ravel()]
    if inel is not None:
        parts.append(np.asarray(inel, dtype=np.float64).ravel())
    if nu is not None:
        parts.append(np.asarray(nu, dtype=np.float64).ravel())
    np.concatenate(parts).astype("<f8").tofile(path)


def read_result(path):
    """(el_mat, inel_mat, nuinel_mat) as [NE][G][L] arrays (None where the reference leaves them unallocated);
    for an S(a,b) case el_mat is scatt_mat."""
    a = np.fromfile(path, dtype="<f8")
    _, ne_el, G, L, ne_in, has_nu = a[:6]
    ne_el, G, L, ne_in = int(ne_el), int(G), int(L), int(ne_in)
    w = G * L
    pos = 6
    el = a[pos:pos + ne_el * w].reshape(ne_el, G, L)
    pos += ne_el * w
    inel = nu = None
    if ne_in:
        inel = a[pos:pos + ne_in * w].reshape(ne_in, G, L)
        pos += ne_in * w
        if has_nu:
            nu = a[pos:pos + ne_in * w].reshape(ne_in, G, L)
            pos += ne_in * w
    assert pos == len(a), "result file has trailing data"
    return el, inel, nu
