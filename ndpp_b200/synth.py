"""Synthetic ACE-shaped nuclides and S(a,b) tables for the BASELINE.json configurations.

No ACE data ships with the reference (data/*.xml only list file names) and there is no network, so
every input is generated here, seeded, directly as the parsed structures of ace.py -- i.e. exactly
what the reference's src/ace.F90 would hand to calc_scatt / calc_scattsab.  Shapes follow SURVEY.md
section 8(d):
  C1  tests/test_scatt fixture (tests/test_scatt/test_scattdata.F90:2183-2275)
  C2  U-238 shape: elastic + 40 discrete levels (law 3) + continuum (law 44, CM), 70 groups
  C3  H-1 free gas at 293.6 / 600 / 1200 K
  C4  H in H2O S(a,b): discrete-skewed and continuous secondary-energy variants
  C5  library of light / medium / heavy nuclides
"""
from __future__ import annotations

import numpy as np

from .ace import (ANGLE_32_EQUI, ANGLE_ISOTROPIC, ANGLE_TABULAR, ELASTIC, K_BOLTZMANN, LINEAR_LINEAR, N_FISSION,
                  N_LEVEL, N_NC, SAB_ELASTIC_DISCRETE, SAB_ELASTIC_EXACT, SAB_SECONDARY_CONT, SAB_SECONDARY_EQUAL,
                  SAB_SECONDARY_SKEWED, DistAngle, DistEnergy, DistEnergySab, Nuclide, Params, Reaction, SAlphaBeta,
                  Tab1)

SEED0 = 20261018
KT_293K = 2.5301e-8
KT_600K = 5.1704e-8
KT_1200K = 1.0341e-7


def group_structure(G: int = 70, e_lo: float = 1.0e-9, e_hi: float = 20.0) -> np.ndarray:
    """E_bins(1) = 0 (required, src/ndpp.F90:229) then G log-spaced edges from e_lo to e_hi."""
    return np.concatenate([[0.0], np.geomspace(e_lo, e_hi, G)])


# --------------------------------------------------------------------------------------------------
# angular-distribution builders (raw ACE AND-block layout, src/ace.F90:905-953)
# --------------------------------------------------------------------------------------------------
def _equi32_edges(a: float) -> np.ndarray:
    k = np.arange(33) / 32.0
    if abs(a) < 1e-12:
        e = -1.0 + 2.0 * k
    else:
        e = np.log(np.exp(-a) + k * (np.exp(a) - np.exp(-a))) / a
    e[0], e[-1] = -1.0, 1.0
    return e


def _tabular_block(a: float, NP: int, interp: int = LINEAR_LINEAR) -> np.ndarray:
    mu = np.linspace(-1.0, 1.0, NP)
    mu[-1] = 1.0
    pdf = np.exp(a * mu)
    area = np.sum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(mu))
    pdf = pdf / area
    cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(mu))])
    return np.concatenate([[float(interp), float(NP)], mu, pdf, cdf])


def make_adist(energies, kinds, aniso, NP_tab=33) -> DistAngle:
    """kinds[i] in {ANGLE_ISOTROPIC, ANGLE_32_EQUI, ANGLE_TABULAR}; aniso[i] = a in f ~ exp(a mu)."""
    data = [np.zeros(1)]  # data(1) is padding so that a location of 0 can mean "isotropic"
    loc = []
    pos = 1
    for k, a in zip(kinds, aniso):
        if k == ANGLE_ISOTROPIC:
            loc.append(0)
            continue
        blk = _equi32_edges(a) if k == ANGLE_32_EQUI else _tabular_block(a, NP_tab)
        loc.append(pos)  # data(lc+1) is the first value of the block
        data.append(blk)
        pos += len(blk)
    return DistAngle(energy=np.asarray(energies, float), type=np.asarray(kinds, np.int32),
                     location=np.asarray(loc, np.int32), data=np.concatenate(data))


def make_law44(e_in, rows) -> np.ndarray:
    """rows[i] = (INTT', Eout, pdf, cdf, R, A); returns edist%data (src/ace.F90:1120-1250 layout)."""
    NE = len(e_in)
    head = 2 + 2 * NE
    data = [np.array([0.0, float(NE)]), np.asarray(e_in, float), np.zeros(NE)]
    pos = head
    locs = []
    for (intt, Eout, pdf, cdf, R, A) in rows:
        locs.append(pos)
        blk = np.concatenate([[float(intt), float(len(Eout))], Eout, pdf, cdf, R, A])
        data.append(blk)
        pos += len(blk)
    data[2] = np.asarray(locs, float)
    return np.concatenate(data)


def _lin_cdf(x, p):
    return np.concatenate([[0.0], np.cumsum(0.5 * (p[1:] + p[:-1]) * np.diff(x))])


# --------------------------------------------------------------------------------------------------
# C1: the reference's own calc_scatt fixture
# --------------------------------------------------------------------------------------------------
def c1_fixture(mt_level=(51, 52)):
    """tests/test_scatt/test_scattdata.F90:2183-2275.  Returns (nuclide, e_bins, params).

    The fixture gives its two inelastic reactions MT = N_LEVEL (4); the *current* is_valid_scatter
    (src/scattdata_header.F90:1508: MT == 2 or 11 <= MT <= 91) rejects MT 4, so with the verbatim
    MTs (mt_level=(4, 4)) only the elastic reaction is integrated.  The default relabels them
    MT 51 / 52 -- every other number is verbatim -- so that the discrete-inelastic and Law-44 paths
    the configuration names are exercised."""
    r1 = Reaction(MT=ELASTIC, Q_value=0.0, multiplicity=1, threshold=1, scatter_in_cm=True, sigma=np.zeros(3))
    ad = DistAngle(energy=np.array([1.5, 2.5]), type=np.array([ANGLE_TABULAR] * 2, np.int32),
                   location=np.array([1, 7], np.int32),
                   data=np.array([0.0, LINEAR_LINEAR, 2.0, -1.0, 1.0, 0.5, 0.5,
                                  LINEAR_LINEAR, 2.0, -1.0, 1.0, 0.5, 0.5, 0.0]))
    r2 = Reaction(MT=mt_level[0], Q_value=0.0, multiplicity=1, threshold=1, scatter_in_cm=True,
                  sigma=np.array([1.0, 0.5, 0.25]), adist=ad)
    ed_data = np.array([0.0, 2.0, 2.0, 3.0, 6.0, 18.0, 2.0, 2.0, 1.0, 2.0, 0.5, 0.5, 0.0, 1.0, 0.5, 0.0, 0.5, 0.5,
                        2.0, 2.0, 2.0, 3.0, 0.5, 0.5, 0.0, 1.0, 1.0, 0.0, 1.0, 0.5, 0.0])
    ed2 = DistEnergy(law=66)
    ed1 = DistEnergy(law=44, data=ed_data, p_valid=Tab1(x=np.array([1.0, 2.0]), y=np.array([0.5, 1.0])), next=ed2)
    r3 = Reaction(MT=mt_level[1], Q_value=-0.02, multiplicity=2, threshold=2, scatter_in_cm=False,
                  sigma=np.array([1.0, 2.0]), edist=ed1)
    r4 = Reaction(MT=N_FISSION)
    nuc = Nuclide(awr=100.0, kT=0.0, energy=np.array([1.0, 2.0, 3.0]), elastic=np.array([0.25, 0.5, 1.0]),
                  reactions=[r1, r2, r3, r4], freegas_cutoff=0.0, name="c1-fixture")
    return nuc, np.array([1.0, 2.0, 3.0]), Params(order=5, mu_bins=3001, nuscatter=True)


def c1_ein_grid(n_extra: int = 97, seed: int = SEED0 + 1) -> np.ndarray:
    """E_in points for C1: the nuclide grid and group edges, the adist/edist break points, seeded
    points in between, and the reference's E_top*(1+1e-3) extra point (src/scatt.F90:426-447).

    The fixture's angular distribution of reaction 2 is tabulated at E = 1.5 and 2.5 only, and the
    current interp_distro looks E_in up with binary_search (src/scattdata_header.F90:471-475), which
    is a fatal error above the last tabulated energy (src/search.F90:36-38) -- so the reference can
    only run this fixture for E_in <= 2.5 (and for the extra point above the top group edge, which
    copies the previous column)."""
    rng = np.random.default_rng(seed)
    pts = np.concatenate([[1.0, 1.5, 2.0, 2.5], rng.uniform(1.0, 2.5, n_extra)])
    pts = np.unique(pts)
    return np.concatenate([pts, [3.0 * (1.0 + np.float32(1.0e-3))]])


# --------------------------------------------------------------------------------------------------
# C2: U-238-shaped heavy nuclide
# --------------------------------------------------------------------------------------------------
def heavy_nuclide(n_grid=20000, n_levels=40, awr=236.0058, kT=KT_293K, seed=SEED0 + 2, with_continuum=True,
                  n_ein_cont=30, np_cont=64, e_max=20.0, name="U-238-shape", first_level=0.0449, level_step=0.028,
                  q_cont=-1.2, n_el_adist=40, n_lvl_adist=20, np_lvl=21) -> Nuclide:
    rng = np.random.default_rng(seed)
    energy = np.unique(np.exp(rng.uniform(np.log(1.0e-11), np.log(e_max), n_grid - 2)))
    energy = np.concatenate([[1.0e-11], energy, [e_max]])
    energy = np.unique(energy)
    n_grid = len(energy)

    # elastic: 9 b potential + 1/v + lognormal resonances
    el = 9.0 + 0.02 / np.sqrt(energy / 2.53e-8)
    e_res = np.exp(rng.uniform(np.log(6.0e-6), np.log(2.0e-2), 200))
    for er, amp, wid in zip(e_res, rng.lognormal(3.0, 1.0, 200), rng.uniform(0.002, 0.02, 200)):
        el += amp / (1.0 + ((energy - er) / (wid * er)) ** 2)

    # elastic angular distribution: isotropic < 10 keV, 32-equiprobable to 100 keV, tabular above
    ead_e = np.geomspace(1.0e-11, e_max, n_el_adist)
    ead_e[0], ead_e[-1] = 1.0e-11, e_max
    kinds = np.where(ead_e < 1.0e-2, ANGLE_ISOTROPIC, np.where(ead_e < 1.0e-1, ANGLE_32_EQUI, ANGLE_TABULAR))
    aniso = 6.0 * np.clip(np.log(ead_e / 1.0e-2) / np.log(e_max / 1.0e-2), 0.0, 1.0)
    rxns = [Reaction(MT=ELASTIC, Q_value=0.0, threshold=1, scatter_in_cm=True, sigma=np.zeros(0),
                     adist=make_adist(ead_e, kinds, aniso))]

    for k in range(n_levels):
        Q = -(first_level + level_step * k)
        e_thr = (awr + 1.0) / awr * abs(Q)
        thr = int(np.searchsorted(energy, e_thr, side="left")) + 1  # 1-based first grid E >= e_thr
        if thr >= n_grid:
            continue
        eg = energy[thr - 1:]
        x = np.log(eg / eg[0] + 1.0e-30)
        sig = 0.5 * rng.uniform(0.3, 1.0) * x * np.exp(1.0 - x / rng.uniform(0.6, 1.4)) / rng.uniform(0.6, 1.4)
        sig = np.maximum(sig, 0.0)
        sig[0] = 0.0
        sig[1:] = np.maximum(sig[1:], 1.0e-6)
        ad_e = np.geomspace(eg[0], e_max, n_lvl_adist)
        ad_e[0], ad_e[-1] = eg[0], e_max
        an = rng.uniform(0.0, 3.0) * np.linspace(0.0, 1.0, n_lvl_adist)
        rxns.append(Reaction(MT=51 + k, Q_value=Q, threshold=thr, scatter_in_cm=True, sigma=sig,
                             adist=make_adist(ad_e, [ANGLE_TABULAR] * n_lvl_adist, an, NP_tab=np_lvl),
                             edist=DistEnergy(law=3, data=np.array([e_thr, (awr / (awr + 1.0)) ** 2]))))

    if with_continuum:
        Q = q_cont
        e_thr = (awr + 1.0) / awr * abs(Q)
        thr = int(np.searchsorted(energy, e_thr, side="left")) + 1
        eg = energy[thr - 1:]
        sig = 2.5 * (1.0 - np.exp(-(eg - eg[0]) / 1.5))
        sig[1:] = np.maximum(sig[1:], 1.0e-6)
        e_in = np.geomspace(eg[0], e_max, n_ein_cont)
        e_in[0], e_in[-1] = eg[0], e_max
        rows = []
        for E in e_in:
            emax = max((E + Q * (awr + 1.0) / awr) * (awr / (awr + 1.0)) ** 2, 1.0e-3)
            Eout = np.linspace(0.0, emax, np_cont)
            T = 0.3 + 0.05 * E
            pdf = Eout * np.exp(-Eout / T)
            pdf[0] = 0.0
            pdf /= np.sum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(Eout))
            cdf = _lin_cdf(Eout, pdf)
            R = 0.9 * (1.0 - Eout / emax)
            A = 0.2 + 3.0 * Eout / emax
            rows.append((2, Eout, pdf, cdf, R, A))
        rxns.append(Reaction(MT=N_NC, Q_value=Q, threshold=thr, scatter_in_cm=True, sigma=sig,
                             edist=DistEnergy(law=44, data=make_law44(e_in, rows),
                                              p_valid=Tab1(x=np.array([eg[0], e_max]), y=np.array([1.0, 1.0])))))
    return Nuclide(awr=awr, kT=kT, energy=energy, elastic=el, reactions=rxns, freegas_cutoff=0.0, name=name)


def c2_u238(n_grid=20000, **kw):
    """Returns (nuclide, e_bins, params, Ein_el, Ein_inel): the nuclide grid is the E_in grid."""
    nuc = heavy_nuclide(n_grid=n_grid, **kw)
    e_bins = group_structure(70)
    params = Params(order=7, mu_bins=2001, nuscatter=False)
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != ELASTIC)
    Ein_el = nuc.energy.copy()
    Ein_inel = nuc.energy[nuc.energy >= thr].copy()
    return nuc, e_bins, params, Ein_el, Ein_inel


# --------------------------------------------------------------------------------------------------
# C3: H-1 free gas
# --------------------------------------------------------------------------------------------------
def c3_h1_freegas(kT=KT_293K, n_ein=1000, cutoff_kT=400.0, order=3, mu_bins=2001):
    energy = np.geomspace(1.0e-11, 20.0, 600)
    energy[0], energy[-1] = 1.0e-11, 20.0
    nuc = Nuclide(awr=0.999167, kT=kT, energy=energy, elastic=np.full(len(energy), 20.4),
                  reactions=[Reaction(MT=ELASTIC, Q_value=0.0, threshold=1, scatter_in_cm=True)],
                  freegas_cutoff=cutoff_kT * kT, name=f"H-1 kT={kT:g}")
    e_bins = group_structure(70)
    Ein = np.geomspace(1.0e-11, cutoff_kT * kT, n_ein)
    return nuc, e_bins, Params(order=order, mu_bins=mu_bins), Ein


# --------------------------------------------------------------------------------------------------
# C4: S(a,b)
# --------------------------------------------------------------------------------------------------
def _equi_cosines(a, n):
    q = (np.arange(n) + 0.5) / n
    a = np.where(np.abs(a) < 1e-9, 1e-9, a)
    return np.log(np.exp(-a) + q * (np.exp(a) - np.exp(-a))) / a


def c4_sab(mode="skewed", n_ein=116, n_eout=64, n_mu=16, kT=KT_293K, elastic=None, seed=SEED0 + 4, e_max=4.0e-6):
    """H-in-H2O-shaped thermal table.  mode: 'equal' | 'skewed' | 'cont'.
    elastic: None | 'coherent' (graphite-like Bragg edges, exact) | 'incoherent' (discrete cosines)."""
    from scipy.stats import gamma
    rng = np.random.default_rng(seed)
    e_in = np.geomspace(1.0e-11, e_max, n_ein)
    sigma = 20.0 + 60.0 / np.sqrt(1.0 + e_in / kT)
    awr = 0.999167

    def eout_quantiles(E, q):
        return gamma.ppf(q, 2.0, scale=0.5 * kT + 0.45 * E)

    def aniso(E, Eo):
        return np.clip(0.8 * np.sqrt(E * Eo) / kT, 0.0, 5.0)

    kw = {}
    if mode in ("equal", "skewed"):
        q = (np.arange(n_eout) + 0.5) / n_eout
        e_out = np.stack([eout_quantiles(E, q) for E in e_in])                    # [n_ein][n_eout]
        mu = np.stack([np.stack([_equi_cosines(aniso(E, Eo), n_mu) for Eo in row])
                       for E, row in zip(e_in, e_out)])                            # [n_ein][n_eout][n_mu]
        kw = dict(inelastic_e_out=e_out, inelastic_mu=mu,
                  secondary_mode=SAB_SECONDARY_EQUAL if mode == "equal" else SAB_SECONDARY_SKEWED)
    else:
        rows = []
        for E in e_in:
            n = int(rng.integers(60, 401)) if n_eout is None or n_eout <= 0 else int(rng.integers(60, max(61, n_eout)))
            sc = 0.5 * kT + 0.45 * E
            Eo = np.concatenate([[0.0], np.geomspace(1.0e-4 * sc, 14.0 * sc, n - 1)])
            pdf = gamma.pdf(Eo, 2.0, scale=sc)
            pdf /= np.sum(pdf[:-1] * np.diff(Eo))
            mu = np.stack([_equi_cosines(aniso(E, x), n_mu) for x in Eo])          # [n][n_mu]
            rows.append(DistEnergySab(e_out=Eo, e_out_pdf=pdf, mu=mu))
        kw = dict(inelastic_data=rows, secondary_mode=SAB_SECONDARY_CONT)

    if elastic == "coherent":
        # Bragg edges; threshold_elastic = last tabulated energy (src/ace.F90:1504)
        edges = np.concatenate([np.cumsum(rng.uniform(0.5e-9, 3.0e-9, 39)) + 1.8e-9, [e_max]])
        kw.update(threshold_elastic=float(edges[-1]), elastic_mode=SAB_ELASTIC_EXACT, elastic_e_in=edges,
                  elastic_P=np.cumsum(rng.uniform(0.2e-9, 2.0e-9, 40)), elastic_mu=None)
    elif elastic == "incoherent":
        ee = np.geomspace(1.0e-11, e_max, 30)
        kw.update(threshold_elastic=float(ee[-1]), elastic_mode=SAB_ELASTIC_DISCRETE, elastic_e_in=ee,
                  elastic_P=5.0 + 20.0 * np.exp(-ee / (20 * kT)),
                  elastic_mu=np.stack([_equi_cosines(min(4.0, E / (4 * kT)), 10) for E in ee]))
    return SAlphaBeta(awr=awr, kT=kT, threshold_inelastic=float(e_in[-1]), inelastic_e_in=e_in,
                      inelastic_sigma=sigma, n_inelastic_mu=n_mu, **kw)


# --------------------------------------------------------------------------------------------------
# C5: library
# --------------------------------------------------------------------------------------------------
def c5_library(n_nuclides=300, seed=SEED0 + 5, ne_lo=2000, ne_hi=40000):
    """Yield (nuclide, Ein_el, Ein_inel) for a library of three shapes (light: elastic only; medium:
    10 levels + continuum; heavy: C2 shape).  Deterministic in (seed, index)."""
    rng = np.random.default_rng(seed)
    specs = []
    for i in range(n_nuclides):
        shape = ("light", "medium", "heavy")[int(rng.integers(0, 3))]
        awr = float(np.exp(rng.uniform(np.log(1.0), np.log(250.0))))
        ne = int(np.exp(rng.uniform(np.log(ne_lo), np.log(ne_hi))))
        specs.append((i, shape, awr, ne))
    return specs


def c5_nuclide(spec, seed=SEED0 + 5):
    i, shape, awr, ne = spec
    nlev = {"light": 0, "medium": 10, "heavy": 40}[shape]
    # lighter targets have higher first levels; keep the thresholds inside the grid
    first = 0.0449 if shape == "heavy" else 0.5
    nuc = heavy_nuclide(n_grid=ne, n_levels=nlev, awr=max(awr, 1.0001), seed=seed * 1000 + i,
                        with_continuum=(shape != "light"), name=f"c5-{i:03d}-{shape}", first_level=first,
                        level_step=0.028 if shape == "heavy" else 0.15, q_cont=-1.2 if shape == "heavy" else -3.0)
    Ein_el = nuc.energy.copy()
    inel = [nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != ELASTIC]
    Ein_inel = nuc.energy[nuc.energy >= min(inel)].copy() if inel else None
    return nuc, Ein_el, Ein_inel


def c5_shape(spec, e_max=20.0):
    """library.NuclideShape of a C5 nuclide from its spec alone (no tables are generated): every rank
    plans the whole library without building the nuclides it does not own."""
    from .library import NuclideShape
    i, shape, awr, ne = spec
    awr = max(awr, 1.0001)
    nlev = {"light": 0, "medium": 10, "heavy": 40}[shape]
    first = 0.0449 if shape == "heavy" else 0.5
    step = 0.028 if shape == "heavy" else 0.15
    q_cont = -1.2 if shape == "heavy" else -3.0
    lev = tuple((awr + 1.0) / awr * (first + step * k) for k in range(nlev))
    cont = (awr + 1.0) / awr * abs(q_cont) if shape != "light" else None
    thr = [x for x in lev] + ([cont] if cont is not None else [])
    n_inel = 0
    if thr:
        frac = np.log(e_max / min(thr)) / np.log(e_max / 1.0e-11)
        n_inel = max(1, int(ne * frac))
    return NuclideShape(index=i, n_el=ne, n_inel=n_inel, level_thresholds=lev, cont_threshold=cont, e_lo=1.0e-11,
                        e_hi=e_max)


# --------------------------------------------------------------------------------------------------
# fissionable nuclides for chi (row N4): synthetic shapes, no evaluated data
# --------------------------------------------------------------------------------------------------
def _watt_rows(e_in, n_out, a0=0.988, b=2.249, e_max=20.0):
    """Law-4 rows with a Watt-shaped spectrum whose temperature drifts with E_in (lin-lin, cdf by trapezoid)."""
    rows = []
    for k, E in enumerate(e_in):
        a = a0 * (1.0 + 0.01 * k)
        Eout = np.concatenate([[0.0], np.geomspace(1.0e-5, e_max, n_out - 1)])
        pdf = np.exp(-Eout / a) * np.sinh(np.sqrt(b * Eout))
        cdf = _lin_cdf(Eout, pdf)
        pdf, cdf = pdf / cdf[-1], cdf / cdf[-1]
        cdf[-1] = 1.0
        rows.append((2, Eout, pdf, cdf, np.zeros(0), np.zeros(0)))
    return rows


def _tab1_data(x, y, interp=None):
    """[NR, (NBT, INT), NE, x, y] as the reference keeps nu / yield / temperature tables in a flat array."""
    head = [0.0] if interp is None else [1.0, float(len(x)), float(interp)]
    return np.concatenate([head, [float(len(x))], np.asarray(x, float), np.asarray(y, float)])


def fissile_total(n_grid=300, n_ein=12, n_out=80, n_precursor=6, seed=SEED0 + 6):
    """U-235 shape: one total-fission reaction (MT 18) with a law-4 spectrum, tabular nu-bar, tabular delayed
    nu-bar, `n_precursor` delayed groups each with its own law-4 spectrum and energy-dependent yield."""
    rng = np.random.default_rng(seed)
    energy = np.geomspace(1.0e-11, 20.0, n_grid)
    energy[0], energy[-1] = 1.0e-11, 20.0
    fis = 1.2 + 580.0 * np.sqrt(2.53e-8 / energy)
    e_in = np.concatenate([[1.0e-11], np.geomspace(1.0e-6, 20.0, n_ein - 1)])
    e_in[-1] = 20.0
    ed = DistEnergy(law=4, data=make_law44(e_in, _watt_rows(e_in, n_out)),
                    p_valid=Tab1(x=np.array([1.0e-11, 20.0]), y=np.array([1.0, 1.0])))
    rxn = Reaction(MT=18, Q_value=193.0, multiplicity=19, threshold=1, scatter_in_cm=False, sigma=fis.copy(), edist=ed)
    nu_e = np.array([1.0e-11, 1.0, 5.0, 20.0])
    prec, dl = [], []
    for k in range(n_precursor):
        ye = np.array([1.0e-11, 4.0, 20.0])
        yk = rng.uniform(0.05, 0.3, 3)
        prec.append(np.concatenate([[0.0125 * 3.0 ** k], _tab1_data(ye, yk)]))
        e_d = np.array([1.0e-11, 20.0]) if k % 2 == 0 else np.array([1.0e-11, 2.0e-6 * (k + 1), 20.0])
        dl.append(DistEnergy(law=4, data=make_law44(e_d, _watt_rows(e_d, 40, a0=0.4 + 0.05 * k, b=1.0, e_max=8.0)),
                             p_valid=Tab1(x=np.array([1.0e-11, 20.0]), y=np.array([1.0, 1.0]))))
    return Nuclide(awr=233.0248, kT=KT_293K, energy=energy, elastic=np.full(n_grid, 11.0),
                   reactions=[Reaction(MT=ELASTIC, threshold=1), rxn], name="U-235-shape",
                   nu_t_type=2, nu_t_data=_tab1_data(nu_e, [2.43, 2.55, 3.10, 5.2]),
                   nu_d_type=2, nu_d_data=_tab1_data(nu_e, [0.0167, 0.0167, 0.0150, 0.0090]),
                   nu_d_precursor_data=np.concatenate(prec), nu_d_edist=dl)


def fissile_partial(n_grid=240, seed=SEED0 + 7):
    """Pu-240 shape: partial fission reactions MT 19 / 20 / 21 / 38 with the analytic laws -- Maxwell (7),
    evaporation (9), Watt (11) -- and a nested chain (law 4 valid with probability p(E), then law 7) whose
    p_valid carries an interpolation region, as the reference's `Pu-240 issue` comments describe
    (src/chi.F90:47-49, src/chidata_header.F90:199,210); polynomial nu-bar, no delayed data.  The restriction
    energies are chosen so that the reference's formulas stay finite: its law 7 and law 11 replace a group edge
    by U itself (src/chidata_header.F90:358,362,431,438), which is NaN for the negative U of evaluated fission
    data whenever the replacement triggers."""
    rng = np.random.default_rng(seed)
    energy = np.geomspace(1.0e-11, 20.0, n_grid)
    energy[0], energy[-1] = 1.0e-11, 20.0
    e_t = np.array([1.0e-11, 1.0, 20.0])
    maxwell = lambda U: np.concatenate([_tab1_data(e_t, [1.29, 1.33, 1.62]), [U]])
    evap = lambda U: np.concatenate([_tab1_data(e_t, [0.9, 1.0, 1.4], interp=2), [U]])
    watt = lambda U: np.concatenate([_tab1_data(e_t, [0.96, 0.98, 1.05]), _tab1_data(e_t, [2.2, 2.25, 2.5]), [U]])
    one = Tab1(x=np.array([1.0e-11, 20.0]), y=np.array([1.0, 1.0]))
    thr20, thr21, thr38 = (int(np.searchsorted(energy, e)) + 1 for e in (5.5, 11.5, 17.0))
    e4 = np.array([1.0e-11, 0.5, 6.0, 20.0])
    chain = DistEnergy(law=4, data=make_law44(e4, _watt_rows(e4, 60)),
                       p_valid=Tab1(x=np.array([1.0e-11, 3.0, 20.0]), y=np.array([1.0, 0.7, 0.4]),
                                    nbt=np.array([3], np.int32), int=np.array([2], np.int32)),
                       next=DistEnergy(law=7, data=maxwell(-20.0),
                                       p_valid=Tab1(x=np.array([1.0e-11, 3.0, 20.0]), y=np.array([0.0, 0.3, 0.6]),
                                                    nbt=np.array([3], np.int32), int=np.array([2], np.int32))))
    sig = lambda thr, s: s * (1.0 + 0.1 * rng.random(n_grid - thr + 1))
    rx = [Reaction(MT=ELASTIC, threshold=1),
          Reaction(MT=19, Q_value=190.0, multiplicity=19, threshold=1, scatter_in_cm=False, sigma=sig(1, 1.5), edist=chain),
          Reaction(MT=20, Q_value=185.0, multiplicity=19, threshold=thr20, scatter_in_cm=False, sigma=sig(thr20, 0.6),
                   edist=DistEnergy(law=9, data=evap(5.4), p_valid=one)),
          Reaction(MT=21, Q_value=180.0, multiplicity=19, threshold=thr21, scatter_in_cm=False, sigma=sig(thr21, 0.4),
                   edist=DistEnergy(law=11, data=watt(2.0), p_valid=one)),
          Reaction(MT=38, Q_value=175.0, multiplicity=19, threshold=thr38, scatter_in_cm=False, sigma=sig(thr38, 0.2),
                   edist=DistEnergy(law=7, data=maxwell(16.0), p_valid=one))]
    return Nuclide(awr=237.9916, kT=KT_293K, energy=energy, elastic=np.full(n_grid, 10.0), reactions=rx, name="Pu-240-shape",
                   nu_t_type=1, nu_t_data=np.array([3.0, 2.8, 0.14, 0.002]))
