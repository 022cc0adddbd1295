import numpy as np

RTOL = 1e-9   # BASELINE.json north_star: 1e-9 relative or 1e-12 absolute per moment
ATOL = 1e-12


def assert_parity(got, ref, rtol=RTOL, atol=ATOL, what=""):
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert np.all(np.isfinite(got) == np.isfinite(ref)), f"{what}: non-finite pattern differs"
    fin = np.isfinite(ref)
    err = np.abs(got[fin] - ref[fin])
    ok = (err <= rtol * np.abs(ref[fin])) | (err <= atol)
    if not np.all(ok):
        k = np.argmax(np.where(ok, 0.0, err))
        raise AssertionError(f"{what}: {np.count_nonzero(~ok)} of {ok.size} moments outside rel {rtol} / abs {atol}; "
                             f"worst |d|={err[k]:.3e} ref={ref[fin][k]:.6e} got={got[fin][k]:.6e}")
    return float(err.max()) if err.size else 0.0


def small_heavy(n_grid=400, n_levels=6, seed=7, **kw):
    """A scaled-down C2 nuclide the oracle integrates in seconds."""
    from ndpp_b200 import synth
    return synth.heavy_nuclide(n_grid=n_grid, n_levels=n_levels, seed=seed, n_ein_cont=8, np_cont=12, n_el_adist=12,
                               n_lvl_adist=6, np_lvl=9, **kw)


def heavy_limit_law61(b=0.6, awr=1.0e8, NP=41, uniform=False):
    """Analytic pin of unit-base interpolation + integrate_file6_cm_leg (src/scattdata_header.F90:1521-1717, 1085-1266),
    for which the reference holds no test: one CM continuum reaction (Law 61) on an infinitely heavy target, so that
    CM = lab (c -> 0, J -> 1), whose tables are separable -- pdf(E_out) = 2 E / Emax^2 on [0, Emax] (Emax = 0.8 at
    E_in = 1, 2.0 at E_in = 3: unit-base interpolation gives the same triangle on the interpolated Emax) times the
    angular density 0.5 (1 + b mu) (P0 = 1, P1 = b/3, higher moments 0).  sigma = p_valid = 1, so the inelastic matrix is
    the normalised distribution itself.  Returns (nuclide, energy_bins, params, Emax(E_in))."""
    from ndpp_b200 import ace
    e_in, emax = np.array([1.0, 3.0]), np.array([0.8, 2.0])
    blocks, locs = [], []
    pos = 2 + 2 * len(e_in)
    for Em in emax:
        Eout = np.linspace(0.0, Em, NP)
        ang = [np.array([2.0, 2.0, -1.0, 1.0, 0.5 * (1 - b), 0.5 * (1 + b), 0.0, 1.0]) for _ in range(NP)]
        LC = pos + 2 + 4 * NP + 8.0 * np.arange(NP)
        locs.append(pos)
        pdf, cdf = (np.full(NP, 1.0 / Em), Eout / Em) if uniform else (2.0 * Eout / Em ** 2, Eout ** 2 / Em ** 2)
        blocks.append(np.concatenate([[2.0, float(NP)], Eout, pdf, cdf, LC] + ang))
        pos += 2 + 4 * NP + 8 * NP
    data = np.concatenate([[0.0, float(len(e_in))], e_in, np.asarray(locs, float)] + blocks)
    energy = np.geomspace(1e-3, 20.0, 60)
    rxn = ace.Reaction(MT=91, Q_value=-0.1, threshold=1, scatter_in_cm=True, sigma=np.ones(len(energy)),
                       edist=ace.DistEnergy(law=61, data=data,
                                            p_valid=ace.Tab1(x=np.array([energy[0], 20.0]), y=np.array([1.0, 1.0]))))
    nuc = ace.Nuclide(awr=awr, kT=0.0, energy=energy, elastic=np.ones(len(energy)),
                      reactions=[ace.Reaction(MT=2, threshold=1), rxn])
    e_bins = np.array([0.0, 0.2, 0.5, 0.9, 1.4, 5.0])
    return nuc, e_bins, ace.Params(order=3, mu_bins=201), (lambda E: np.interp(E, e_in, emax))


def assert_heavy_limit(m, e_bins, Emax, b=0.6):
    """m[g][l] at one E_in against the closed forms of heavy_limit_law61."""
    assert abs(m[:, 0].sum() - 1.0) < 1e-12                      # normalised to sum_g P0 = 1 (:1255-1264)
    edges = np.minimum(e_bins, Emax)
    p0 = (edges[1:] ** 2 - edges[:-1] ** 2) / Emax ** 2
    inside = np.nonzero(e_bins[1:] < Emax)[0]                    # groups wholly below the kinematic edge
    assert len(inside) >= 2
    # the trapezoid over the NE_PER_GRP outgoing energies is exact for the linear pdf, so interior groups are in the
    # analytic ratio; the group that holds Emax loses part of its last interval (the integrand drops to zero within
    # 1e-8 of Emax), an inherent feature of the reference's quadrature: checked to 2 %
    assert np.allclose(m[inside, 0] / m[inside[0], 0], p0[inside] / p0[inside[0]], rtol=1e-6)
    assert np.allclose(m[:, 0], p0, atol=0.02)
    assert np.allclose(m[inside, 1] / m[inside, 0], b / 3.0, rtol=1e-6)
    assert np.all(np.abs(m[inside, 2:] / m[inside, :1]) < 1e-6)
    above = np.nonzero(e_bins[:-1] >= Emax)[0]
    assert np.all(np.abs(m[above, 0]) < 1e-6)                    # nothing beyond the unit-base interpolated Emax


def freegas_a1_analytic_p0(E, kT, e_bins):
    """Group probabilities of the free-gas scattering kernel of a target of mass ratio A = 1 with a constant,
    isotropic-in-CM cross section (the closed form of e.g. Bell & Glasstone, Nuclear Reactor Theory, section 7.3:
    sigma(E -> E') E = erf(sqrt(E'/kT)) for E' < E and exp((E - E')/kT) erf(sqrt(E/kT)) for E' > E, with
    sigma_s(E) = (1 + kT/2E) erf(sqrt(E/kT)) + exp(-E/kT) / sqrt(pi E/kT)), integrated over every group with scipy."""
    from scipy import integrate, special
    x = E / kT

    def kern(xp):
        return special.erf(np.sqrt(xp)) if xp < x else np.exp(x - xp) * special.erf(np.sqrt(x))
    sig = (1.0 + 0.5 / x) * special.erf(np.sqrt(x)) + np.exp(-x) / np.sqrt(np.pi * x)
    p = np.zeros(len(e_bins) - 1)
    for g in range(len(p)):
        lo, hi = e_bins[g] / kT, e_bins[g + 1] / kT
        p[g] = integrate.quad(kern, lo, hi, points=[x] if lo < x < hi else None, epsabs=1e-14, epsrel=1e-12,
                              limit=400)[0] / (x * sig)
    return p


def freegas_analytic_p0(E, kT, e_bins, A):
    """Group probabilities of the free-gas kernel for a target of mass ratio A (constant cross section, isotropic in
    CM; Bell & Glasstone section 7.3): with eta = (A+1)/(2 sqrt A), rho = (A-1)/(2 sqrt A), x = E/kT, x' = E'/kT,
      sigma(E -> E') ~ eta^2/(2x) { erf(eta sqrt x' - rho sqrt x) +- erf(eta sqrt x' + rho sqrt x)
                                    + exp(x - x') [erf(eta sqrt x - rho sqrt x') -+ erf(eta sqrt x + rho sqrt x')] }
    (upper signs for E' < E), normalised by sigma_s(E) ~ [(b^2 + 1/2) erf b + b exp(-b^2)/sqrt pi] / b^2, b^2 = A x."""
    import warnings
    from scipy import integrate, special
    eta, rho = (A + 1.0) / (2.0 * np.sqrt(A)), (A - 1.0) / (2.0 * np.sqrt(A))
    x = E / kT

    def kern(xp):
        a, b = np.sqrt(xp), np.sqrt(x)
        s = 1.0 if xp < x else -1.0
        return 0.5 * eta * eta / x * (special.erf(eta * a - rho * b) + s * special.erf(eta * a + rho * b) +
                                      np.exp(x - xp) * (special.erf(eta * b - rho * a) - s * special.erf(eta * b + rho * a)))
    b2 = A * x
    tot = ((b2 + 0.5) * special.erf(np.sqrt(b2)) + np.sqrt(b2 / np.pi) * np.exp(-b2)) / b2
    p = np.zeros(len(e_bins) - 1)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")          # the erf differences cancel far from E; quad says so
        for g in range(len(p)):
            lo, hi = e_bins[g] / kT, e_bins[g + 1] / kT
            p[g] = integrate.quad(kern, lo, hi, points=[x] if lo < x < hi else None, epsabs=1e-14, epsrel=1e-12,
                                  limit=400)[0] / tot
    return p


def check_library_against_oracle(n_check):
    """`keep` callback of ndpp_b200.library.run_c5: sampled nuclides of a finished library run, assembled on the root
    device (ndppgpu_library_fetch), against the CPU oracle at 1e-9 rel / 1e-12 abs; the counts go into the report."""
    def keep(out, run, specs, parsed, e_bins, params):
        import os

        from ndpp_b200 import synth
        from oracle import pyoracle
        rng = np.random.default_rng(5)
        worst = {"nuclides": [], "cells": 0, "outside_1e-9rel_1e-12abs": 0, "max_abs": 0.0}
        for i in [s[0] for s in specs][:: max(1, len(specs) // n_check)][:n_check]:
            nuc, Eel, Einel = parsed[i] if i in parsed else synth.c5_nuclide(specs[i])
            rn = pyoracle.RefNuclide(nuc, e_bins, params)
            rn.convert_distro()
            for m, E in ((0, Eel), (1, Einel)):
                if E is None or len(E) == 0:
                    continue
                mat = run.fetch(i, m, E, e_bins[-1])
                idx = np.sort(rng.choice(np.nonzero(E <= e_bins[-1])[0], min(24, len(E)), replace=False))
                ref = rn.elastic(E[idx], n_threads=os.cpu_count()) if m == 0 else \
                    rn.inelastic(E[idx], n_threads=os.cpu_count())[0]
                err = np.abs(mat[idx] - ref)
                worst["cells"] += int(err.size)
                worst["outside_1e-9rel_1e-12abs"] += int(((err > RTOL * np.abs(ref)) & (err > ATOL)).sum())
                worst["max_abs"] = max(worst["max_abs"], float(err.max()))
            worst["nuclides"].append(int(i))
            rn.close()
        out["parity_vs_oracle"] = worst
    return keep


def check_against_rest_vectors(sab_calc, inelastic_of, thin_grid, apply_tol):
    """An implementation (the oracle on the CPU, the CUDA path on the GPU) against tests/golden/walk_vectors_rest.npz: the
    committed results of the independent numpy evaluations of tests/walks.py (scripts/make_walk_golden_rest.py).
      sab_calc(sab, e_bins, order, E, parts) -> out | (out, el, inel);  inelastic_of(nuc, e_bins, params, E) -> [NE][G][L]
      thin_grid(x, y, tokeep, tol) -> kept indices;  apply_tol(d, tol) -> array"""
    import os

    from tests import walks
    v = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "walk_vectors_rest.npz"))
    for mode in ("equal", "skewed"):
        sab, e_bins, E = walks.sab_discrete_case(mode)
        assert np.array_equal(E, v[f"sab_disc_{mode}_Ein"])
        assert np.allclose(sab_calc(sab, e_bins, 5, E, False), v[f"sab_disc_{mode}"], rtol=1e-11, atol=1e-13), mode
    for elastic in ("coherent", "incoherent"):
        sab, e_bins, E = walks.sab_elastic_case(elastic)
        out, el, inel = sab_calc(sab, e_bins, 5, E, True)
        ref, sig = v[f"sab_el_{elastic}"], v[f"sab_el_{elastic}_sig"]
        for i in range(len(E) - 1):
            assert np.allclose(el[i], ref[i], rtol=1e-11, atol=1e-13 * sig[i]), (elastic, i)
            tot = el[i] + inel[i]
            assert np.allclose(out[i], tot / tot[:, 0].sum(), rtol=1e-12, atol=1e-15)
    sab, e_bins, E = walks.sab_continuous_case()
    out, el, inel = sab_calc(sab, e_bins, 5, E, True)
    assert np.allclose(inel[:-1], v["sab_cont_inel"][:-1], rtol=1e-10, atol=1e-12 * np.max(sab.inelastic_sigma))
    assert np.allclose(out, v["sab_cont"], rtol=1e-10, atol=1e-13)
    nuc, e_bins, params, spec, Ein = walks.law9_case()
    m = inelastic_of(nuc, e_bins, params, Ein)
    p = v["law9_p0"]
    assert np.allclose(m[:, :, 0], p, rtol=1e-9, atol=1e-13)
    nz = p > 1e-12
    assert np.allclose((m[:, :, 1] / np.where(nz, m[:, :, 0], 1.0))[nz], walks.LAW9_B / 3.0, rtol=1e-8)
    assert np.all(np.abs(m[:, :, 2:][nz] / m[:, :, :1][nz]) < 1e-8)
    nuc, e_bins, params, Ein = walks.file6_lab_case()
    m = inelastic_of(nuc, e_bins, params, Ein)[0]
    assert np.allclose(m[:, 0], v["file6_lab_p0"], rtol=0, atol=1e-12)
    assert np.allclose(m[:, 1] / m[:, 0], float(v["file6_lab_p1_over_p0"]), rtol=1e-9) and np.all(np.abs(m[:, 2:]) < 1e-9)
    x, y, tokeep, tol = walks.thin_case()
    assert np.array_equal(thin_grid(x, y, tokeep, tol), v["thin_keep"])
    d, tol = walks.tol_case()
    got = apply_tol(d, tol)
    assert np.allclose(got, v["tol_out"], rtol=1e-15, atol=0.0) and np.all(got[9] == 0.0)
