"""Evaluations of the parity-unpinned routines that are independent of the C restatement (oracle/) and of CUDA: written
from the Fortran text with numpy (numpy's Legendre polynomials, its own searches, scipy quadrature), each with the case
it runs on.  tests/test_oracle_golden.py holds the oracle against them, scripts/make_walk_golden_rest.py commits their
results as tests/golden/walk_vectors_rest.npz, and tests/test_gpu_parity.py compares the CUDA path with the committed
vectors directly -- so that for these routines the GPU evidence does not pass only through the restatement.

    routine                              reference                         case / function
    integrate_sab_inel_disc + combine    src/sab.F90:142-245, 415-454      sab_discrete_case / walk_sab_discrete
    integrate_sab_el + combine           src/sab.F90:21-109, 415-454       sab_elastic_case / walk_sab_elastic
    integrate_sab_inel_cont              src/sab.F90:253-408               sab_continuous_case / walk_sab_continuous
    law9_scatter_lab_leg                 scattdata_header.F90:1274-1326    law9_case / walk_law9
    integrate_file6_lab_leg              scattdata_header.F90:1334-1450    file6_lab_case / walk_file6_lab
    thin_grid                            src/thin.F90:51-169               thin_case / walk_thin_grid
    apply_tol_scatt                      src/scatt.F90:786-818             tol_case / walk_apply_tol
"""
import numpy as np
from numpy.polynomial import legendre as npleg

from ndpp_b200 import ace, synth


def _bsearch(a, v):
    """0-based lower index, v == last -> n - 2 (src/search.F90:21-71)"""
    return min(int(np.searchsorted(a, v, side="right")) - 1, len(a) - 2)


# ---- S(a,b), discrete secondary energies ---------------------------------------------------------------------------------
def sab_discrete_case(mode):
    sab = synth.c4_sab(mode=mode, elastic=None, n_ein=20, n_eout=16, n_mu=8)
    e_bins = synth.group_structure(30, 1e-10, 1e-5)
    rng = np.random.default_rng(8)
    ein = np.asarray(sab.inelastic_e_in)
    E = np.sort(np.concatenate([ein[[3, 11]], np.exp(rng.uniform(np.log(ein[0]), np.log(ein[-1]), 12))]))
    return sab, e_bins, E


def walk_sab_discrete(sab, e_bins, E, mode, L=6):
    """Combined matrix without an elastic part: sum_{E_out in g} w_j sum_k P_l(mu_jk) over the interpolated table row,
    divided by its P0 total; the last column copies its predecessor (src/sab.F90:452)."""
    ein = np.asarray(sab.inelastic_e_in)
    eo, mu = np.asarray(sab.inelastic_e_out), np.asarray(sab.inelastic_mu)    # [iEin][iEout], [iEin][iEout][imu]
    n_out, n_mu = eo.shape[1], mu.shape[2]
    w = np.ones(n_out)
    if mode == "skewed":
        w[[0, -1]], w[[1, -2]] = 0.1, 0.4
    w = w / (w.sum() * n_mu)
    out = np.zeros((len(E), len(e_bins) - 1, L))
    for i, e in enumerate(E[:-1]):
        k = min(int(np.searchsorted(ein, e, side="right")) - 1, len(ein) - 2)
        f = (e - ein[k]) / (ein[k + 1] - ein[k])
        eo_i = (1 - f) * eo[k] + f * eo[k + 1]
        mu_i = (1 - f) * mu[k] + f * mu[k + 1]
        ref = out[i]
        for j in range(n_out):
            if e_bins[0] <= eo_i[j] < e_bins[-1]:
                g = int(np.searchsorted(e_bins, eo_i[j], side="right")) - 1
                for l in range(L):
                    ref[g, l] += w[j] * npleg.legval(mu_i[j], [0] * l + [1]).sum()
        ref /= ref[:, 0].sum()
    out[-1] = out[-2]
    return out


# ---- S(a,b), elastic part and combination -----------------------------------------------------------------------------------
def sab_elastic_case(elastic):
    sab = synth.c4_sab(mode="equal", elastic=elastic, n_ein=20, n_eout=16, n_mu=8)
    e_bins = synth.group_structure(30, 1e-10, 1e-5)
    rng = np.random.default_rng(9)
    ee = np.asarray(sab.elastic_e_in)
    E = np.sort(np.exp(rng.uniform(np.log(ee[0] * 1.01), np.log(ee[-1] * 0.99), 15)))
    return sab, e_bins, E


def walk_sab_elastic(sab, e_bins, E, elastic, L=6):
    """integrate_sab_el: coherent = one cosine 1 - E_bragg / E weighted P / E, incoherent = equally likely interpolated
    cosines weighted by the interpolated P.  Returns the elastic partial integrals (rows of E[:-1])."""
    ee, P = np.asarray(sab.elastic_e_in), np.asarray(sab.elastic_P)
    el = np.zeros((len(E), len(e_bins) - 1, L))
    scale = np.zeros(len(E))
    for i, e in enumerate(E[:-1]):
        k = int(np.searchsorted(ee, e, side="right")) - 1
        f = (e - ee[k]) / (ee[k + 1] - ee[k])
        g = int(np.searchsorted(e_bins, e, side="right")) - 1
        if elastic == "coherent":
            mu = np.array([1.0 - ee[k] / e])
            sig, w = P[k] / e, 1.0
        else:
            em = np.asarray(sab.elastic_mu)
            mu = (1 - f) * em[k] + f * em[k + 1]
            sig, w = (1 - f) * P[k] + f * P[k + 1], 1.0 / em.shape[1]
        for l in range(L):
            el[i, g, l] = sig * w * npleg.legval(mu, [0] * l + [1]).sum()
        scale[i] = sig
    return el, scale


# ---- S(a,b), continuous secondary energies ----------------------------------------------------------------------------------
def sab_continuous_case():
    sab = synth.c4_sab(mode="cont", elastic=None, n_ein=10, n_eout=70, n_mu=6)
    e_bins = synth.group_structure(24, 1e-10, 1e-5)
    ein = np.asarray(sab.inelastic_e_in)
    rng = np.random.default_rng(10)
    E = np.sort(np.concatenate([ein[[2]], np.exp(rng.uniform(np.log(ein[0]), np.log(ein[-1] * 0.999), 10))]))
    return sab, e_bins, E


def walk_sab_continuous(sab, e_bins, E, L=6):
    """Stage 1 integrates every table row over the groups with the weights pdf(i) * dE(i) and the edge rule of the text
    (f * bin at both edges, cosines interpolated to the edge), stage 2 interpolates linearly to E_in and scales with the
    interpolated sigma.  Returns (inelastic partial integrals, combined matrix) for the rows of E[:-1]."""
    G = len(e_bins) - 1
    ein, sg = np.asarray(sab.inelastic_e_in), np.asarray(sab.inelastic_sigma)

    def pl(mu):                      # sum over the cosines of P_0..P_{L-1}
        return np.array([npleg.legval(mu, [0] * l + [1]).sum() for l in range(L)])

    rows = []
    for d in sab.inelastic_data:
        Eo, mu = np.asarray(d.e_out), np.asarray(d.mu)        # mu[iEout][imu]
        w = np.append(np.asarray(d.e_out_pdf)[:-1] * np.diff(Eo), 0.0)
        dist = np.zeros((G, L))
        for g in range(G):
            lo_e, hi_e = e_bins[g], e_bins[g + 1]
            acc = np.zeros(L)
            if lo_e < Eo[0]:
                i_lo = 0
            elif lo_e >= Eo[-1]:
                continue
            else:
                i = _bsearch(Eo, lo_e)
                f = (lo_e - Eo[i]) / (Eo[i + 1] - Eo[i])
                acc += f * w[i] * pl((1 - f) * mu[i] + f * mu[i + 1])
                i_lo = i + 1
            if hi_e < Eo[0]:
                continue
            elif hi_e >= Eo[-1]:
                i_hi = len(Eo) - 2
            else:
                i = _bsearch(Eo, hi_e)
                f = (hi_e - Eo[i]) / (Eo[i + 1] - Eo[i])
                acc += f * w[i] * pl((1 - f) * mu[i] + f * mu[i + 1])
                i_hi = i - 1
            for i in range(i_lo, i_hi + 1):
                acc += w[i] * pl(mu[i])
            dist[g] = acc / mu.shape[1]
        rows.append(dist)
    rows = np.array(rows)
    inel = np.zeros((len(E), G, L))
    out = np.zeros_like(inel)
    for i, e in enumerate(E[:-1]):
        k = _bsearch(ein, e)
        f = (e - ein[k]) / (ein[k + 1] - ein[k])
        inel[i] = ((1 - f) * rows[k] + f * rows[k + 1]) * ((1 - f) * sg[k] + f * sg[k + 1])
        out[i] = inel[i] / inel[i][:, 0].sum()
    out[-1] = out[-2]
    return inel, out


# ---- law 9 (evaporation spectrum, laboratory angular table) ----------------------------------------------------------------
LAW9_B = 0.45


def law9_case():
    b = LAW9_B
    energy = np.geomspace(1e-11, 20.0, 80)
    thr = int(np.searchsorted(energy, 1.0)) + 1
    e0 = energy[thr - 1]
    e9, T9, U = np.array([e0, 5.0, 20.0]), np.array([0.3, 0.6, 1.1]), 0.4
    d9 = np.concatenate([[0.0, 3.0], e9, T9, [U]])
    blk = np.array([2.0, 2.0, -1.0, 1.0, 0.5 * (1 - b), 0.5 * (1 + b), 0.0, 1.0])      # lin-lin, 2 points
    # the angular table is read on the energy grid of the law-9 block (scattdata_header.F90:342-368), so it has its rows
    ad = ace.DistAngle(energy=e9.copy(), type=np.array([ace.ANGLE_TABULAR] * 3, np.int32),
                       location=np.array([1, 9, 17], np.int32), data=np.concatenate([[0.0], blk, blk, blk]))
    pv = ace.Tab1(x=np.array([e0, 20.0]), y=np.array([1.0, 1.0]))
    r9 = ace.Reaction(MT=16, Q_value=-0.9, threshold=thr, scatter_in_cm=False, multiplicity=1,
                      sigma=np.ones(len(energy) - thr + 1), adist=ad, edist=ace.DistEnergy(law=9, data=d9, p_valid=pv))
    nuc = ace.Nuclide(awr=26.7, kT=0.0, energy=energy, elastic=np.full(len(energy), 2.0),
                      reactions=[ace.Reaction(MT=2, threshold=1), r9])
    e_bins = synth.group_structure(20, 1e-4, 20.0)
    return nuc, e_bins, ace.Params(order=4, mu_bins=401), (e9, T9, U), np.array([1.7, 4.0, 12.0])


def walk_law9(e_bins, spec, Ein):
    """Group probabilities of E' exp(-E'/T) up to E - U by numerical quadrature (sigma = p_valid = 1: the inelastic column
    is the distribution itself); the angular moments of a linear table are P1/P0 = b/3, higher moments 0."""
    from scipy import integrate
    e9, T9, U = spec
    T = float(np.interp(Ein, e9, T9))
    top = Ein - U
    norm = integrate.quad(lambda e: e * np.exp(-e / T), 0.0, top, epsabs=0, epsrel=1e-13)[0]
    return np.array([integrate.quad(lambda e: e * np.exp(-e / T), min(lo, top), min(hi, top), epsabs=0, epsrel=1e-13)[0]
                     for lo, hi in zip(e_bins[:-1], e_bins[1:])]) / norm


# ---- integrate_file6_lab_leg ----------------------------------------------------------------------------------------------------
def file6_lab_case():
    from tests.util import heavy_limit_law61
    nuc, e_bins, params, _ = heavy_limit_law61(awr=55.0, uniform=True)
    nuc.reactions[1].scatter_in_cm = False
    return nuc, e_bins, params, np.array([3.0])


def walk_file6_lab():
    """Hand evaluation of the Fortran text on a uniform pdf (41 E_out points on [0, 2], every pdf(i) * dE(i) = 0.025) and
    the group edges (0, 0.2, 0.5, 0.9, 1.4, 5): a lower edge adds f_lo * bin (the part *below* the edge, :1385-1391) and
    then starts at the next bin, an upper edge adds f_hi * bin; edges that coincide with an E_out point have f = 0, and
    linspace puts E_out(29) just above 1.4 (f = 1 on bin 28 for both neighbours).  Bins counted per group: 3, 5, 7, 9,
    13; the final normalisation (:1447-1448) divides by their sum, 37.  Angular part: P1/P0 = b/3 = 0.2."""
    assert np.linspace(0.0, 2.0, 41)[28] > 1.4
    return np.array([3.0, 5.0, 7.0, 9.0, 13.0]) / 37.0, 0.2


# ---- thin_grid, apply_tol_scatt ---------------------------------------------------------------------------------------------------
def thin_case():
    rng = np.random.default_rng(21)
    x = np.geomspace(1e-6, 20.0, 400)
    y = (np.sin(2.0 * np.log(x))[:, None] + 0.3) * np.linspace(1.0, 2.0, 10)[None, :]      # changes sign
    y += 1e-4 * rng.normal(size=y.shape)
    y[50:60] = 0.0                                                                         # y == 0: absolute error
    return x, y, np.array([x[123], 7.0]), 5e-3


def walk_thin_grid(x, y, tokeep, tol):
    """thin_grid_one walked literally: point k is tested against (last kept, k + 1) with log-x interpolation, the error is
    divided by y *with its sign* (src/thin.F90:125-127: a negative y makes any error acceptable), points in `tokeep`
    stay.  Returns the kept indices."""
    ref = [0]
    klo, k = 0, 1
    while k + 1 < len(x):
        frac = 1.0 / np.log(x[k + 1] / x[klo]) * np.log(x[k] / x[klo])
        removable = not np.any(tokeep == x[k])
        if removable:
            t = y[klo] + (y[k + 1] - y[klo]) * frac
            err = np.abs(t - y[k])
            nz = y[k] != 0.0
            err[nz] = err[nz] / y[k][nz]
            removable = bool(np.all(err <= tol))
        if not removable:
            ref.append(k)
            klo = k
        k += 1
    ref.append(len(x) - 1)
    return np.array(ref)


def tol_case():
    rng = np.random.default_rng(22)
    d = rng.normal(size=(60, 11, 5)) * 0.2
    d[:, :, 0] = np.abs(d[:, :, 0]) * (rng.uniform(size=(60, 11)) > 0.2)
    d[::4, 3, 0] = 4e-9
    d[7] = 0.0
    d[9, :, 0] = 0.0                      # orig_total = 0 with non-zero higher moments: norm = 0 wipes the column
    return d, 1e-8


def walk_apply_tol(d, tol):
    """Groups with 0 < P0 < tol are zeroed for every order, then the column is scaled by orig_total / new_total (0 when
    orig_total <= 0)."""
    ref = d.copy()
    for i in range(len(ref)):
        orig = ref[i, :, 0].sum()
        small = (ref[i, :, 0] > 0.0) & (ref[i, :, 0] < tol)
        ref[i, small, :] = 0.0
        ref[i] *= (orig / ref[i, :, 0].sum()) if orig > 0.0 else 0.0
    return ref
