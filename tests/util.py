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


def assert_parity_floor(got, ref, what="", rtol=RTOL, floor=1e-8):
    """Parity at the reference's own round-off floor, for paths whose tables went through libm
    transcendentals (Law 44 sinh/cosh): calc_int_pn_tablelin's closed forms carry a round-off of
    up to ~3e-9 of P0 per moment at the default mu spacing (measured against mpmath), and
    a last-bit change of any table value re-draws it.  |d| <= rtol*|ref| + floor*P0(E_in)."""
    got, ref = np.asarray(got), np.asarray(ref)
    p0 = np.abs(ref[:, :, 0]).sum(axis=1)[:, None, None]
    err = np.abs(got - ref)
    ok = err <= rtol * np.abs(ref) + floor * p0 + ATOL
    assert np.all(ok), f"{what}: {np.count_nonzero(~ok)} of {ok.size} moments outside the round-off floor; " \
                       f"worst |d|/P0 = {np.max(err / np.maximum(p0, 1e-300)):.3e}"
    return float(np.max(err / np.maximum(p0, 1e-300)))


def small_heavy(n_grid=400, n_levels=6, seed=7, **kw):
    """A scaled-down C2 nuclide the oracle integrates in seconds."""
    from ndpp_b200 import synth
    return synth.heavy_nuclide(n_grid=n_grid, n_levels=n_levels, seed=seed, n_ein_cont=8, np_cont=12, n_el_adist=12,
                               n_lvl_adist=6, np_lvl=9, **kw)
