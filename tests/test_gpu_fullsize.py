"""GPU tests at BASELINE.json's full sizes (configs[1] C2 and configs[2] C3), where the oracle cannot integrate the
whole workload in seconds: size-independent properties over every column, a seeded sample of columns against the
oracle, determinism of the dynamically scheduled kernels, and the degenerate sizes the C-ABI must accept."""
import numpy as np
import pytest

from ndpp_b200 import ace, synth
from tests.util import assert_parity, small_heavy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scatt():
    from ndpp_b200 import scatt as s
    s.default_context()
    return s


@pytest.fixture(scope="module")
def c2(scatt):
    nuc, e_bins, params, Ein_el, Ein_inel = synth.c2_u238()
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    el = dn.elastic(Ein_el)
    inel, _ = dn.inelastic(Ein_inel)
    yield dict(nuc=nuc, e_bins=e_bins, params=params, Ein_el=Ein_el, Ein_inel=Ein_inel, dn=dn, el=el, inel=inel)
    dn.clear()


def test_c2_full_size_properties(c2):
    el, inel, nuc = c2["el"], c2["inel"], c2["nuc"]
    assert el.shape == (20000, 70, 8) and inel.shape == (len(c2["Ein_inel"]), 70, 8) and inel.shape[0] > 4000
    assert np.isfinite(el).all() and np.isfinite(inel).all()
    # target-at-rest elastic: a probability distribution over the groups per E_in.  The angular tables are
    # integrated by the trapezoid rule on the uniform mu grid, so the sum is 1 up to that rule's O(dmu^2) error
    p0 = el[:, :, 0]
    assert p0.min() >= 0.0 and np.allclose(p0.sum(axis=1), 1.0, atol=1.5e-4)      # measured on B200: 1.4e-5
    # |P_l| <= P_0 for a non-negative density (|P_l(mu)| <= 1), group by group
    assert np.all(np.abs(el) <= p0[:, :, None] * (1 + 1e-9) + 1e-12)
    assert np.all(np.abs(inel) <= inel[:, :, :1] * (1 + 1e-9) + 1e-10)
    # no up-scatter off a target at rest: nothing above the group that holds E_in
    g_in = np.searchsorted(c2["e_bins"], c2["Ein_el"], side="right") - 1
    for k in range(0, 20000, 97):
        assert np.all(p0[k, min(g_in[k], 69) + 1:] == 0.0)
    # inelastic matrices are weighted by the reaction cross sections (multiplicity 1, p_valid 1): the P0 sum of a
    # column is the sum of the open reactions' cross sections at E_in (the grid is the nuclide grid: no interpolation)
    sig = np.zeros(len(nuc.energy))
    for r in nuc.reactions[1:]:
        sig[r.threshold - 1:r.threshold - 1 + len(r.sigma)] += r.sigma
    off = len(nuc.energy) - inel.shape[0]
    tot = inel[:, :, 0].sum(axis=1)
    assert np.allclose(tot[:-1], sig[off:-1], rtol=1e-14, atol=1e-14)              # measured: 6.6e-16 relative


def test_c2_full_size_sample_against_oracle(c2, oracle):
    rng = np.random.default_rng(2)
    rn = oracle.RefNuclide(c2["nuc"], c2["e_bins"], c2["params"])
    rn.convert_distro()
    dn = c2["dn"]
    ke = np.sort(rng.choice(20000, 32, replace=False))
    assert_parity(c2["el"][ke], rn.elastic(c2["Ein_el"][ke]), what="C2 elastic sample")
    # the Law 44 tables converted on the device are the oracle's bit for bit (csrc/libm_exact.cuh)
    n = 0
    for s in range(dn.n_slots):
        info = dn.slot_info(s)
        if info["is_init"] and info["law"] == 44:
            for iE in range(1, info["NE"] + 1):
                assert np.array_equal(dn.get_table(s, iE)[0], rn.get_table(s, iE)[0]), (s, iE)
                n += 1
    assert n == 30
    ki = np.sort(rng.choice(len(c2["Ein_inel"]), 16, replace=False))
    gi, _ = dn.inelastic(c2["Ein_inel"][ki])
    ri, _ = rn.inelastic(c2["Ein_inel"][ki])
    assert np.any(ri != 0)
    assert_parity(gi, ri, what="C2 inelastic sample")
    rn.close()


def test_dynamic_scheduling_is_deterministic(scatt):
    """The file-6 pipeline and the free-gas kernel pull tasks from global counters; the moments must not depend on
    which warp took which task."""
    nuc, e_bins, params, Ein_el, Ein_inel = synth.c2_u238(n_grid=3000)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    a, _ = dn.inelastic(Ein_inel)
    b, _ = dn.inelastic(Ein_inel)
    assert np.array_equal(a, b)
    dn.clear()
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=60)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    a, b = dn.elastic(Ein), dn.elastic(Ein)
    assert np.array_equal(a, b)
    dn.clear()


def test_c3_full_size_properties_and_sample(scatt, oracle):
    nuc, e_bins, params, Ein = synth.c3_h1_freegas()
    assert len(Ein) == 1000
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    el = dn.elastic(Ein)
    assert el.shape == (1000, 70, 4) and np.isfinite(el).all()
    assert np.allclose(el[:, :, 0].sum(axis=1), 1.0, atol=5e-15)       # integrate_freegas_leg normalises (:131-140); measured 4.4e-16
    # every order has its own adaptivity (tolerances 1e-7 / 1e-8): measured max(|P_l| - P0) = 2.4e-7
    assert np.all(np.abs(el) <= np.abs(el[:, :, :1]) + 2.5e-6)
    # up-scatter exists below a few kT and has died out at the cutoff
    g_in = np.searchsorted(e_bins, Ein, side="right") - 1
    up = np.array([el[k, g_in[k] + 1:, 0].sum() for k in range(1000)])
    assert up[:300].min() > 1e-3 and up[-1] < 0.05
    rn = oracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    k = np.array([3, 377, 640, 998])
    assert_parity(el[k], rn.elastic(Ein[k]), what="C3 sample")
    rn.close()
    dn.clear()


@pytest.mark.parametrize("order", [0, 10])
def test_extreme_orders_and_degenerate_sizes(scatt, oracle, order):
    """Order 0 and MAX_LEGENDRE_ORDER = 10 (src/constants.F90:113; l = 9 repeats the l = 7 closed form,
    src/legendre.F90:117-126), one group, one E_in, an empty E_in grid."""
    nuc = small_heavy(n_grid=200, n_levels=3)
    e_bins = np.array([0.0, 20.0]) if order == 0 else synth.group_structure(9, 1e-7, 20.0)
    params = ace.Params(order=order, mu_bins=301, nuscatter=True)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    rn = oracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    Ein = np.array([1e-9, 0.3, 2.0, 7.5, 20.0])
    assert_parity(dn.elastic(Ein), rn.elastic(Ein), what=f"elastic order {order}")
    gi, gn = dn.inelastic(Ein[2:])
    ri, rnu = rn.inelastic(Ein[2:])
    assert np.any(ri != 0)
    assert_parity(gi, ri, what=f"inelastic order {order}")
    assert_parity(gn, rnu, what=f"nu-inelastic order {order}")
    one = dn.elastic(Ein[1:2])
    assert one.shape == (1, len(e_bins) - 1, order + 1) and np.array_equal(one[0], dn.elastic(Ein)[1])
    assert dn.elastic(np.zeros(0)).shape == (0, len(e_bins) - 1, order + 1)
    rn.close()
    dn.clear()
