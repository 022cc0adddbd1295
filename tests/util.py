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
