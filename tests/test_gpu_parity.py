"""GPU parity tests: the CUDA path, called through the C-ABI (libndppgpu.so), against the CPU
oracle on the same seeded inputs.  Tolerance (BASELINE.json): 1e-9 relative or 1e-12 absolute per
moment; table conversion without transcendentals is bit-exact."""
import numpy as np
import pytest

from ndpp_b200 import ace, synth
from tests.util import assert_parity, small_heavy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def scatt():
    from ndpp_b200 import scatt as s
    s.default_context()  # raises without a GPU / library: no fallback
    return s


def _pair(scatt, oracle, nuc, e_bins, params):
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    rn = oracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    return dn, rn


def test_c1_tables_bit_exact(scatt, oracle):
    nuc, e_bins, params = synth.c1_fixture()
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    assert dn.n_slots == rn.n_slots == 5
    for s in range(5):
        a, b = dn.slot_info(s), rn.slot_info(s)
        assert a["is_init"] == b["is_init"]
        if not a["is_init"]:
            continue
        assert (a["NE"], a["law"], a["has_adist"], a["has_edist"], a["order"]) == \
               (b["NE"], b["law"], b["has_adist"], b["has_edist"], b["order"])
        for iE in range(1, a["NE"] + 1):
            ta, tb = dn.get_table(s, iE), rn.get_table(s, iE)
            # Law 44 included: the device's sinh / cosh carry the host libm's bits (csrc/libm_exact.cuh)
            for k in (0, 1, 2, 3):
                assert np.array_equal(ta[k], tb[k]), (s, iE, k)
            assert ta[4] == tb[4]


def test_c1_moments(scatt, oracle):
    nuc, e_bins, params = synth.c1_fixture()
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    Ein = synth.c1_ein_grid()
    assert_parity(dn.elastic(Ein), rn.elastic(Ein), what="C1 elastic")
    ri, rnu = rn.inelastic(Ein)
    gi, gn = dn.inelastic(Ein)          # device-converted Law 44 tables, the default path
    assert_parity(gi, ri, what="C1 inelastic")
    assert_parity(gn, rnu, what="C1 nu-inelastic")
    # the extra point above the top group edge copies the previous column (src/scatt.F90:669,770)
    assert np.array_equal(gi[-1], gi[-2]) and np.any(gi[-2] != 0)


def test_c1_calc_scatt_signature(scatt, oracle):
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(20)
    el, inel, nu = scatt.calc_scatt(nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, True, Ein, Ein[Ein >= 2.0])
    assert el.shape == (len(Ein), 2, 6) and inel.shape[0] == np.count_nonzero(Ein >= 2.0) and nu is not None
    el2, inel2, nu2 = scatt.calc_scatt(nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, False, Ein, None)
    assert inel2 is None and nu2 is None and np.array_equal(el, el2)


@pytest.mark.parametrize("awr", [236.0058, 11.9, 0.999167])
def test_heavy_shape_tables_and_moments(scatt, oracle, awr):
    nuc = small_heavy(awr=awr, first_level=0.0449 if awr > 100 else 0.5, level_step=0.05)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7, mu_bins=2001, nuscatter=True)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    for s in range(dn.n_slots):
        info = dn.slot_info(s)
        assert info == rn.slot_info(s)
        for iE in (1, info["NE"]):
            ta, tb = dn.get_table(s, iE), rn.get_table(s, iE)
            assert np.array_equal(ta[0], tb[0]), (s, iE, info["law"])
    rng = np.random.default_rng(3)
    Eel = np.sort(np.concatenate([nuc.energy[::7], np.exp(rng.uniform(np.log(1e-11), np.log(20.0), 60)), [20.0]]))
    assert_parity(dn.elastic(Eel), rn.elastic(Eel), what=f"elastic awr={awr}")
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != 2)
    Einel = np.sort(np.concatenate([nuc.energy[nuc.energy >= thr][::9], rng.uniform(thr, 20.0, 25), [20.0]]))
    ri, rnu = rn.inelastic(Einel)
    assert np.any(ri != 0)
    gi, gn = dn.inelastic(Einel)        # device-converted Law 44 tables, the default path
    assert_parity(gi, ri, what=f"inelastic awr={awr}")
    assert_parity(gn, rnu, what=f"nu-inelastic awr={awr}")


def _law61_nuclide(lab, intt=2):
    """Continuum reaction with Law 61 tables (one isotropic, tabular hist and lin-lin columns)."""
    rng = np.random.default_rng(61)
    energy = np.geomspace(1e-11, 20.0, 120)
    e_in = np.array([2.0, 5.0, 11.0, 20.0])
    blocks, locs = [], []
    head = 2 + 2 * len(e_in)
    pos = head
    for E in e_in:
        NP = int(rng.integers(5, 9))
        Eout = np.linspace(0.0, 0.6 * E, NP)
        pdf = np.exp(-Eout / (0.2 * E)); pdf /= np.sum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(Eout))
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(Eout))])
        row_len = 2 + 4 * NP
        ang, LC = [], []
        apos = pos + row_len
        for j in range(NP):
            if j % 3 == 0:
                LC.append(0.0)
                continue
            npa = int(rng.integers(3, 12))
            mu = np.linspace(-1, 1, npa); mu[-1] = 1.0
            p = np.exp(rng.uniform(0, 2) * mu); p /= np.sum(0.5 * (p[1:] + p[:-1]) * np.diff(mu))
            c = np.concatenate([[0.0], np.cumsum(0.5 * (p[1:] + p[:-1]) * np.diff(mu))])
            LC.append(float(apos))
            blk = np.concatenate([[float(1 + (j % 2)), float(npa)], mu, p, c])
            ang.append(blk); apos += len(blk)
        locs.append(pos)
        blocks.append(np.concatenate([[float(intt), float(NP)], Eout, pdf, cdf, LC] + ang))
        pos = apos
    data = np.concatenate([[0.0, float(len(e_in))], e_in, np.asarray(locs, float)] + blocks)
    thr = int(np.searchsorted(energy, 2.0)) + 1
    e_in[0] = energy[thr - 1]
    data[2] = e_in[0]
    sig = np.linspace(0.0, 1.5, len(energy) - thr + 1); sig[1:] += 0.1
    rxn = ace.Reaction(MT=91, Q_value=-1.9, threshold=thr, scatter_in_cm=not lab, sigma=sig,
                       edist=ace.DistEnergy(law=61, data=data, p_valid=ace.Tab1(x=np.array([e_in[0], 20.0]),
                                                                                y=np.array([0.8, 1.0]))))
    el = ace.Reaction(MT=2, threshold=1)
    return ace.Nuclide(awr=55.3, kT=0.0, energy=energy, elastic=np.full(len(energy), 3.0), reactions=[el, rxn])


@pytest.mark.parametrize("lab", [False, True])
def test_law61_cm_and_lab(scatt, oracle, lab):
    nuc = _law61_nuclide(lab)
    e_bins = synth.group_structure(30, 1e-5, 20.0)
    params = ace.Params(order=5, mu_bins=501)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    for iE in range(1, 5):
        ta, tb = dn.get_table(1, iE), rn.get_table(1, iE)
        assert np.array_equal(ta[0], tb[0]) and np.array_equal(ta[1], tb[1]) and ta[4] == tb[4]
    Ein = np.array([2.5, 3.0, 4.999, 5.0, 7.7, 11.0, 15.0, 19.0, 20.0])
    gi, _ = dn.inelastic(Ein)
    ri, _ = rn.inelastic(Ein)
    assert np.any(ri != 0)
    assert_parity(gi, ri, what=f"law 61 lab={lab}")


def test_law9_and_law4_lab(scatt, oracle):
    energy = np.geomspace(1e-11, 20.0, 80)
    thr = int(np.searchsorted(energy, 1.0)) + 1
    sig = np.full(len(energy) - thr + 1, 0.7)
    # law 9 evaporation spectrum: [NR=0, NE, E(NE), T(NE), U]
    e9 = np.array([energy[thr - 1], 5.0, 20.0])
    d9 = np.concatenate([[0.0, 3.0], e9, [0.3, 0.6, 1.1], [0.4]])
    ad = synth.make_adist([energy[thr - 1], 8.0, 20.0], [ace.ANGLE_TABULAR] * 3, [0.0, 0.8, 1.5], NP_tab=9)
    pv = ace.Tab1(x=np.array([energy[thr - 1], 20.0]), y=np.array([1.0, 1.0]))
    r9 = ace.Reaction(MT=16, Q_value=-0.9, threshold=thr, scatter_in_cm=False, multiplicity=2, sigma=sig, adist=ad,
                      edist=ace.DistEnergy(law=9, data=d9, p_valid=pv))
    # law 4 (tabular E_out, lab) with an angular distribution
    rows = []
    e4 = np.array([energy[thr - 1], 6.0, 20.0])
    for E in e4:
        Eout = np.linspace(0.0, 0.5 * E, 7)
        pdf = np.ones(7) / (0.5 * E)
        rows.append((2, Eout, pdf, np.linspace(0, 1, 7), np.zeros(0), np.zeros(0)))
    d4 = synth.make_law44(e4, rows)
    r4 = ace.Reaction(MT=22, Q_value=-0.9, threshold=thr, scatter_in_cm=False, sigma=sig,
                      adist=synth.make_adist([energy[thr - 1], 20.0], [ace.ANGLE_32_EQUI] * 2, [0.5, 1.0]),
                      edist=ace.DistEnergy(law=4, data=d4, p_valid=pv))
    nuc = ace.Nuclide(awr=26.7, kT=0.0, energy=energy, elastic=np.full(len(energy), 2.0),
                      reactions=[ace.Reaction(MT=2, threshold=1), r9, r4])
    e_bins = synth.group_structure(20, 1e-4, 20.0)
    params = ace.Params(order=4, mu_bins=401, nuscatter=True)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    Ein = np.array([1.2, 3.3, 5.0, 9.0, 20.0])
    gi, gn = dn.inelastic(Ein)
    ri, rnu = rn.inelastic(Ein)
    assert np.any(ri != 0)
    assert_parity(gi, ri, what="law 9 + law 4 lab")
    assert_parity(gn, rnu, what="law 9 + law 4 lab (nu)")


@pytest.mark.parametrize("kT", [synth.KT_293K, synth.KT_600K, synth.KT_1200K])
def test_c3_freegas(scatt, oracle, kT):
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(kT=kT, n_ein=1000)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    sub = Ein[[0, 333, 700, 940, 999]]  # the last point equals the cutoff: target-at-rest branch
    got, ref = dn.elastic(sub), rn.elastic(sub)
    assert np.allclose(ref[:, :, 0].sum(axis=1), 1.0, atol=1e-12)
    # the adaptive integrator takes its accept / split decisions on exp(): the device evaluates it with the host libm's
    # bits (csrc/libm_exact.cuh), walks every order's own tree and keeps its association, so not a cell may leave the
    # tolerance (measured: max |d| 8e-17 with libdevice's exp, before the exact one)
    err = np.abs(got - ref)
    ok = (err <= 1e-9 * np.abs(ref)) | (err <= 1e-12)
    assert np.count_nonzero(~ok) == 0, f"{np.count_nonzero(~ok)} cells outside tolerance"
    assert err.max() < 1e-13


def test_freegas_heavy_target_two_rows(scatt, oracle):
    # A = 15.9 with a non-isotropic CM distribution: exercises both table rows and the Brent clipping
    energy = np.geomspace(1e-11, 20.0, 200)
    ad = synth.make_adist([1e-11, 1e-6, 20.0], [ace.ANGLE_TABULAR] * 3, [0.0, 0.3, 2.0], NP_tab=11)
    nuc = ace.Nuclide(awr=15.858, kT=synth.KT_600K, energy=energy, elastic=np.full(200, 3.8),
                      reactions=[ace.Reaction(MT=2, threshold=1, adist=ad)], freegas_cutoff=400 * synth.KT_600K)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=3, mu_bins=2001)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    sub = np.array([3e-9, 4.1e-7, 1.9e-5])
    got, ref = dn.elastic(sub), rn.elastic(sub)
    err = np.abs(got - ref)
    ok = (err <= 1e-9 * np.abs(ref)) | (err <= 1e-12)
    assert np.count_nonzero(~ok) == 0
    assert err.max() < 1e-13


@pytest.mark.parametrize("mode,elastic", [("skewed", None), ("equal", "coherent"), ("cont", "incoherent")])
def test_c4_sab(scatt, oracle, mode, elastic):
    sab = synth.c4_sab(mode=mode, elastic=elastic, n_ein=40)
    e_bins = synth.group_structure(70)
    rng = np.random.default_rng(4)
    Ein = np.sort(np.concatenate([sab.inelastic_e_in, np.exp(rng.uniform(np.log(1e-11), np.log(4e-6), 300)),
                                  [5e-6, 1e-5]]))
    ds = scatt.DeviceSab(sab)
    got, gel, ginel = ds.calc(e_bins, ace.SCATT_TYPE_LEGENDRE, 5, Ein, parts=True)
    ref, rel, rinel = oracle.sab_calc(sab, e_bins, 5, Ein, parts=True)
    assert_parity(gel, rel, what=f"sab elastic {mode}")
    assert_parity(ginel, rinel, what=f"sab inelastic {mode}")
    assert_parity(got, ref, what=f"sab combined {mode}")
    assert np.array_equal(got[-1], got[-2])
    assert np.array_equal(scatt.calc_scattsab(sab, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 2001, Ein), got)


def test_legendre_leaf_bit_exact(scatt, oracle):
    """calc_pn / calc_int_pn_tablelin on the device are bit-identical to the oracle (K7)."""
    ctx = scatt.default_context()
    rng = np.random.default_rng(0)
    for dx in (1e-3, 6.7e-4, 0.05):
        n = 3000
        xl = rng.uniform(-1, 1 - dx, n); xh = xl + dx
        fl = rng.uniform(0, 2, n); fh = fl + rng.uniform(-.01, .01, n)
        gi, gp = ctx.test_legendre(11, xl, xh, fl, fh)
        ri = np.array([oracle.calc_int_pn_tablelin(11, a, b, c, d) for a, b, c, d in zip(xl, xh, fl, fh)])
        rp = np.array([[oracle.calc_pn(l, a) for l in range(11)] for a in xl])
        assert np.array_equal(gi, ri) and np.array_equal(gp, rp)
    gi, _ = ctx.test_legendre(6, [0.3], [0.3 + 1e-15], [1.0], [2.0])
    assert np.all(gi == 0.0)  # src/legendre.F90:44


def test_interp_distro_per_slot(scatt, oracle):
    nuc = small_heavy()
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7, mu_bins=2001)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != 2)
    E = np.sort(np.random.default_rng(1).uniform(thr, 20, 10))
    for s in range(1, dn.n_slots):
        got = dn.interp_distro(s, E)
        ref = np.array([rn.interp_distro(s, x) for x in E])
        assert_parity(got, ref, what=f"interp_distro slot {s}")


def test_errors_are_loud(scatt):
    from ndpp_b200.capi import NdppGpuError
    nuc, e_bins, params = synth.c1_fixture()
    with pytest.raises(NdppGpuError):
        scatt.DeviceNuclide(nuc, e_bins, ace.Params(order=11))          # above MAX_LEGENDRE_ORDER
    dn = scatt.DeviceNuclide(nuc, e_bins, params, convert=False)
    with pytest.raises(NdppGpuError):
        dn.elastic(np.array([1.5]))                                     # convert_distro not called
    dn.convert_distro()
    with pytest.raises(NdppGpuError, match="binary search"):
        dn.inelastic(np.array([2.7]))   # above the last tabulated adist energy: the reference aborts (search.F90:36-38)
    assert np.all(np.isfinite(dn.inelastic(np.array([2.2]))[0]))       # the latch was cleared


@pytest.mark.parametrize("order", [7, 5])
def test_file6_cm_ws_bit_identical_to_one_role_kernel(scatt, order, monkeypatch):
    """The warp-specialised producer/consumer kernel (k_file6_cm_ws) re-organises the work of
    k_file6_cm, not its arithmetic: both must give the same bits on the Law-44 continuum."""
    from ndpp_b200.capi import Context
    nuc = small_heavy(n_grid=600)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=order)
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT == ace.N_NC)
    Ein = nuc.energy[nuc.energy >= thr][::3]
    outs = []
    for legacy in ("1", "0"):
        monkeypatch.setenv("NDPPGPU_F6_LEGACY", legacy)
        ctx = Context(-1)
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        slot = [s for s in range(dn.n_slots) if dn.slot_info(s)["is_init"] and dn.slot_info(s)["law"] == 44][0]
        outs.append(dn.interp_distro(slot, Ein))
        dn.clear()
    assert np.any(outs[0] != 0.0)
    assert np.array_equal(outs[0], outs[1])


def test_exact_math_sequences(scatt):
    """FastDiv (nvcc's own division sequence with the reciprocal refinement hoisted) must reproduce
    `/` bit for bit, zero dividends and distant binades included."""
    r = scatt.default_context().test_exact_math(seed=20261018, per_thread=4000)
    assert r["pairs"] > 5e8 and r["mismatch"] == 0, r


def test_against_committed_golden_vectors(scatt):
    """CUDA path vs tests/golden/oracle_vectors.npz (oracle outputs committed with scripts/make_golden.py):
    strict 1e-9 / 1e-12 everywhere, the Law 44 paths included (device-converted tables)."""
    import os
    from ndpp_b200 import egrid
    v = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_vectors.npz"))
    nuc, e_bins, params = synth.c1_fixture()
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    assert_parity(dn.elastic(v["c1_Ein"]), v["c1_elastic"], what="golden C1 elastic")
    gi, gn = dn.inelastic(v["c1_Ein"])
    assert_parity(gi, v["c1_inelastic"], what="golden C1 inelastic")
    assert_parity(gn, v["c1_nu_inelastic"], what="golden C1 nu-inelastic")
    dn.clear()
    dn = scatt.DeviceNuclide(small_heavy(), synth.group_structure(70), ace.Params(order=7))
    assert_parity(dn.elastic(v["heavy_Eel"]), v["heavy_elastic"], what="golden heavy elastic")
    assert_parity(dn.inelastic(v["heavy_Ein"])[0], v["heavy_inelastic"], what="golden heavy inelastic")
    dn.clear()
    nuc, e_bins, params, _ = synth.c3_h1_freegas(n_ein=1000)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    assert_parity(dn.elastic(v["freegas_Ein"]), v["freegas_elastic"], what="golden free gas")
    dn.clear()
    e_bins = synth.group_structure(70)
    for mode, kw in (("skewed", {}), ("cont", {"elastic": "incoherent"})):
        sab = synth.c4_sab(mode, **kw)
        got = scatt.DeviceSab(sab).calc(e_bins, ace.SCATT_TYPE_LEGENDRE, 5, v[f"sab_{mode}_Ein"])
        assert_parity(got[:-1], v[f"sab_{mode}"][:-1], what=f"golden S(a,b) {mode}")
    integ, _ = scatt.default_context().test_legendre(8, *v["leaf_args"])
    assert np.array_equal(integ, v["leaf_integrals"])


@pytest.mark.parametrize("mode,kw", [("skewed", {}), ("equal", {"elastic": "coherent"}), ("cont", {"elastic": "incoherent"})])
def test_c4_sab_tabular_histogram(scatt, oracle, mode, kw):
    """C4's 16-bin cosine histogram (scatt_type = TABULAR; semantics defined by this project because
    the reference leaves it unimplemented, parity unpinned): CUDA vs the oracle, and the bins of a group
    sum to the group's P0 of the Legendre output."""
    from ndpp_b200 import egrid
    sab = synth.c4_sab(mode, **kw)
    e_bins = synth.group_structure(70)
    E = egrid.sab_egrid(sab, e_bins)
    ds = scatt.DeviceSab(sab)
    got = ds.calc(e_bins, ace.SCATT_TYPE_TABULAR, 16, E)
    leg = ds.calc(e_bins, ace.SCATT_TYPE_LEGENDRE, 5, E)
    assert got.shape == (len(E), 70, 16)
    assert np.allclose(got.sum(axis=2), leg[:, :, 0], atol=1e-13)
    idx = np.unique(np.concatenate([np.arange(0, len(E), 23), [len(E) - 2, len(E) - 1]]))
    ref = oracle.sab_calc(sab, e_bins, 16, E[idx], tabular=True)
    assert_parity(got[idx][:-1], ref[:-1], what=f"S(a,b) histogram {mode}")
    with pytest.raises(Exception):
        scatt.DeviceNuclide(small_heavy(), e_bins, ace.Params(order=16, scatt_type=ace.SCATT_TYPE_TABULAR))


def test_apply_tol_and_thin_grid_match_oracle(scatt, oracle):
    """N2: the steps after the integrator (apply_tol_scatt, thin_grid) on the device vs the oracle on
    real moment matrices: identical zero pattern / kept points, values bit-identical (apply_tol) or
    identical columns (thin_grid copies)."""
    nuc = small_heavy(n_grid=900)
    e_bins = synth.group_structure(70)
    dn = scatt.DeviceNuclide(nuc, e_bins, ace.Params(order=5))
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != ace.ELASTIC)
    Eel, Einel = nuc.energy.copy(), nuc.energy[nuc.energy >= thr].copy()
    el = dn.elastic(Eel)
    inel, _ = dn.inelastic(Einel)
    for mat in (el, inel):
        ref = oracle.apply_tol_scatt(mat, 1e-8)
        got = scatt.apply_tol_scatt(mat.copy(), 1e-8)
        assert np.array_equal(got == 0.0, ref == 0.0)
        assert np.array_equal(got, ref)
    el_t = oracle.apply_tol_scatt(el, 1e-8)
    inel_t = oracle.apply_tol_scatt(inel, 1e-8)
    for x, y, y2 in ((Eel, el_t, None), (Einel, inel_t, inel_t * 1.5)):
        keep, comp, _, mabs = oracle.thin_grid(x, y, e_bins, 1e-3, y2)
        gx, gy, gy2, gcomp, gmabs = scatt.thin_grid(x, y, e_bins, 1e-3, y2)
        assert len(gx) == len(keep) and np.array_equal(gx, x[keep])
        assert np.array_equal(gy, y[keep]) and gcomp == comp
        if y2 is not None:
            assert np.array_equal(gy2, y2[keep])
        assert abs(gmabs - mabs) <= 1e-12 * max(mabs, 1e-300) + 1e-18
        assert 0 < len(keep) < len(x)


def test_driver_pipeline_matches_oracle_steps(scatt, oracle, tmp_path):
    """src/ndpp.F90:560-702 for one nuclide: integrate, apply the printing tolerance, thin, write the
    library -- through the fused device entry points -- against the same steps composed from the oracle."""
    from ndpp_b200 import driver, output
    nuc = small_heavy(n_grid=700)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, nuscatter=True)
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != ace.ELASTIC)
    Eel, Einel = nuc.energy.copy(), nuc.energy[nuc.energy >= thr].copy()
    path = str(tmp_path / "nuc.bin")
    res = driver.preprocess_nuclide(nuc, e_bins, params, print_tol=1e-8, thin_tol=2e-3, Ein_el=Eel, Ein_inel=Einel,
                                    library_file=path)
    dn = scatt.DeviceNuclide(nuc, e_bins, params)       # the integration itself is covered elsewhere
    el = oracle.apply_tol_scatt(dn.elastic(Eel), 1e-8)
    inel, nu = dn.inelastic(Einel)
    inel, nu = oracle.apply_tol_scatt(inel, 1e-8), oracle.apply_tol_scatt(nu, 1e-8)
    ke = oracle.thin_grid(Eel, el, e_bins, 2e-3)[0]
    ki = oracle.thin_grid(Einel, inel, e_bins, 2e-3, nu)[0]
    assert np.array_equal(res.Ein_el, Eel[ke]) and np.array_equal(res.el_mat, el[ke])
    assert np.array_equal(res.Ein_inel, Einel[ki]) and np.array_equal(res.inel_mat, inel[ki])
    assert np.array_equal(res.nuinel_mat, nu[ki])
    assert 0 < len(ke) < len(Eel) and res.thin_compr_el == (len(Eel) - len(ke)) / len(Eel)
    lib = output.read_library(path)
    assert lib["trailing_bytes"] == 0 and lib["nuscatter"] and np.array_equal(lib["Ein_inel"], Einel[ki])
    assert np.array_equal(lib["grp_index_el"], output.group_index(Eel[ke], e_bins))
    # the file stores the window between the first and last group of positive P0: what lies outside is zero
    assert np.array_equal(lib["elastic"][:, :, 0] > 0, res.el_mat[:, :, 0] > 0)
    assert np.allclose(lib["elastic"][lib["elastic"] != 0], res.el_mat[lib["elastic"] != 0], rtol=0, atol=0)


def test_freegas_scratch_overflow_reruns_with_worst_case_sizes(scatt, monkeypatch):
    """The level-parallel inner integral runs with a capped scratch first; an overflow must be detected
    and the pass repeated with the worst-case sizes, giving the same bits."""
    from ndpp_b200.capi import Context
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=1000)
    Ein = Ein[[300, 800]]
    outs = []
    for cap in ("4096", "8"):
        monkeypatch.setenv("NDPPGPU_FG_CAP", cap)
        ctx = Context(-1)
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        outs.append(dn.elastic(Ein))
        launches = ctx.stats()["launches"]
        dn.clear()
        outs.append(launches)
    assert np.array_equal(outs[0], outs[2]) and np.any(outs[0] != 0)
    assert outs[3] > outs[1]             # the generations were launched again


def test_freegas_work_items_do_not_change_the_bits(scatt, monkeypatch):
    """The outer recursion is cut into work items of `split` levels that are processed generation by generation and
    re-assembled by their postfix programs: the moments must not depend on the cut, and a queue that is too small
    must be detected and the pass repeated with a larger one."""
    from ndpp_b200.capi import Context
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=1000)
    Ein = Ein[[0, 450, 900]]
    outs, items = [], []
    for split, queue in (("3", None), ("2", None), ("1", None), ("0", None), ("2", "512")):
        monkeypatch.setenv("NDPPGPU_FG_SPLIT", split)
        if queue is None:
            monkeypatch.delenv("NDPPGPU_FG_QUEUE", raising=False)
        else:
            monkeypatch.setenv("NDPPGPU_FG_QUEUE", queue)
        ctx = Context(-1)
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        outs.append(dn.elastic(Ein))
        items.append(ctx.stats()["freegas_items"])
        dn.clear()
    assert np.any(outs[0] != 0)
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)
    assert items[0] < items[1] < items[2] < items[3] and items[4] == items[1]


def test_calc_scatt_in_one_call_equals_the_two_calls(scatt):
    """ndppgpu_calc_scatt (elastic + inelastic, the elastic matrices copied to the host while the inelastic kernels run)
    against ndppgpu_elastic + ndppgpu_inelastic: the same bits, with and without nu-scatter, with and without an
    inelastic grid, into caller-owned arrays."""
    nuc, e_bins, params = synth.c1_fixture()
    params = ace.Params(**{**params.__dict__, "nuscatter": True})
    Ein = synth.c1_ein_grid(60)
    Einel = Ein[Ein >= 2.0]
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    try:
        el, (inel, nu) = dn.elastic(Ein), dn.inelastic(Einel, True)
        a, b, c = dn.calc(Ein, Einel, True)
        assert np.array_equal(a, el) and np.array_equal(b, inel) and np.array_equal(c, nu)
        out_el, out_in = np.full_like(el, -1.0), np.full_like(inel, -1.0)
        a, b, c = dn.calc(Ein, Einel, False, el_out=out_el, inel_out=out_in)
        assert a is out_el and b is out_in and c is None
        assert np.array_equal(out_el, el) and np.array_equal(out_in, inel)
        a, b, c = dn.calc(Ein, None)
        assert np.array_equal(a, el) and b is None and c is None
    finally:
        dn.clear()


def test_freegas_chunked_walk_does_not_change_the_bits(scatt, monkeypatch):
    """The inner (mu) recursion is walked level by level, or 128 intervals of a level at a time with the forest below a
    chunk folded before the next chunk starts (NDPPGPU_FG_CHUNK; csrc/kernels_freegas.cuh: fg_warp_simpson_mu): same intervals,
    same value tree, so the same bits -- here on the three heaviest kinds of columns (E_in far below, near and above kT),
    with a small first-attempt scratch so that the retry with the worst-case scratch runs through the chunked walk too."""
    from ndpp_b200.capi import Context
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=1000)
    Ein = Ein[[0, 450, 900]]
    outs = []
    for chunk, cap in ((None, None), ("128", None), ("128", "64")):
        for name, val in (("NDPPGPU_FG_CHUNK", chunk), ("NDPPGPU_FG_CAP", cap)):
            if val is None:
                monkeypatch.delenv(name, raising=False)
            else:
                monkeypatch.setenv(name, val)
        ctx = Context(-1)
        dn = scatt.DeviceNuclide(nuc, e_bins, params, ctx)
        outs.append(dn.elastic(Ein))
        dn.clear()
    assert np.any(outs[0] != 0)
    for o in outs[1:]:
        assert np.array_equal(outs[0], o)


def test_unitbase_and_file6_cm_leg_heavy_target_limit(scatt, oracle):
    """The CUDA path against closed forms where the reference holds no test (unit-base interpolation +
    integrate_file6_cm_leg, A -> infinity, separable tables; tests/util.py: heavy_limit_law61), and against the oracle."""
    from tests.util import assert_heavy_limit, heavy_limit_law61
    nuc, e_bins, params, emax = heavy_limit_law61()
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    Ein = np.array([1.0, 1.3, 2.0, 2.5, 3.0])
    got, _ = dn.inelastic(Ein)
    ref, _ = rn.inelastic(Ein)
    assert_parity(got, ref, what="heavy-target Law 61")
    for i, E in enumerate(Ein):
        assert_heavy_limit(got[i], e_bins, float(emax(E)))


def test_freegas_p0_matches_the_analytic_kernel_for_A1(scatt):
    """The CUDA free-gas path against the closed-form A = 1 kernel (tests/util.py: freegas_a1_analytic_p0)."""
    from tests.util import freegas_a1_analytic_p0
    nuc, eb, params, _ = synth.c3_h1_freegas(n_ein=8)
    nuc.awr = 1.0
    dn = scatt.DeviceNuclide(nuc, eb, params)
    xs = np.array([0.3, 2.0, 10.0, 60.0])
    got = dn.elastic(xs * nuc.kT)
    for i, x in enumerate(xs):
        p = freegas_a1_analytic_p0(x * nuc.kT, nuc.kT, eb)
        assert np.abs(got[i, :, 0] - p).max() < 5e-6, x


def test_cuda_against_the_committed_walk_vectors(scatt):
    """The CUDA path against tests/golden/walk_vectors.npz: moments computed by literal walks of the Fortran text
    (scripts/make_walk_golden.py) with neither the oracle's integrators nor CUDA -- free gas at the reference's default
    adaptive tolerances, and the Law-44
    continuum through unit-base interpolation + integrate_file6_cm_leg (device-converted tables)."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("make_walk_golden", os.path.join(root, "scripts", "make_walk_golden.py"))
    mk = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mk)
    v = np.load(os.path.join(root, "tests", "golden", "walk_vectors.npz"))
    nuc, e_bins, params, Ein = mk.freegas_case()
    got = scatt.DeviceNuclide(nuc, e_bins, params).elastic(Ein)
    ref = v["freegas_moments"]
    err = np.abs(got - ref)
    ok = (err <= 1e-9 * np.abs(ref)) | (err <= 1e-12)
    assert np.count_nonzero(~ok) == 0 and err.max() < 1e-13
    nuc, e_bins, params = mk.file6_case()
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    for E, ref in zip(v["file6_Ein"], v["file6_moments"]):
        m = dn.interp_distro(int(v["file6_slot"]), np.array([float(E)]))[0]
        m = m / m[:, 0].sum()
        # the walk integrates the mu segments by Gauss-Legendre quadrature instead of the closed forms, whose
        # round-off (up to ~3e-9 of the segment's P0 at this mu spacing) is the reference's own: 1e-8 of P0 = 1
        assert np.all(np.abs(m - ref) <= 1e-9 * np.abs(ref) + 1e-8), float(E)


def test_cuda_against_the_committed_vectors_of_the_other_unpinned_routines(scatt):
    """The CUDA path against tests/golden/walk_vectors_rest.npz: independent numpy evaluations (tests/walks.py) of the
    routines for which the reference holds no test -- S(a,b) elastic / discrete / continuous + combine_sab_grid, law 9,
    integrate_file6_lab_leg, thin_grid, apply_tol_scatt -- so that their GPU evidence does not pass through the C
    restatement."""
    from tests.util import check_against_rest_vectors

    def sab_calc(sab, e_bins, order, E, parts):
        ds = scatt.DeviceSab(sab)
        out = ds.calc(e_bins, ace.SCATT_TYPE_LEGENDRE, order, E, parts=parts)
        ds.clear()
        return out

    def inelastic_of(nuc, e_bins, params, E):
        dn = scatt.DeviceNuclide(nuc, e_bins, params)
        out = dn.inelastic(np.asarray(E, dtype=float))[0]
        dn.clear()
        return out

    def thin(x, y, tokeep, tol):
        xk = scatt.thin_grid(x, y, tokeep, tol)[0]
        return np.searchsorted(x, xk)
    check_against_rest_vectors(sab_calc, inelastic_of, thin, lambda d, tol: scatt.apply_tol_scatt(d.copy(), tol))


@pytest.mark.parametrize("index", [0, 1, 2, 7, 11])
def test_c5_sampled_nuclides_of_every_shape(scatt, oracle, index):
    """BASELINE configs[4]: nuclides of the synthetic 300-nuclide library, one or two of each shape (light: elastic only;
    medium: 10 levels + Law-44 continuum; heavy: C2 shape; awr from 1.3 to 236), P5, 70 groups, on sampled E_in of their
    own grids against the oracle -- device-converted tables, strict tolerance."""
    spec = synth.c5_library(300)[index]
    nuc, Eel, Einel = synth.c5_nuclide(spec)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, mu_bins=2001)
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    rng = np.random.default_rng(500 + index)
    ke = np.sort(rng.choice(len(Eel), 24, replace=False))
    assert_parity(dn.elastic(Eel[ke]), rn.elastic(Eel[ke]), what=f"C5 nuclide {index} ({spec[1]}) elastic")
    if Einel is not None:
        ki = np.sort(rng.choice(len(Einel), 16, replace=False))
        ri = rn.inelastic(Einel[ki])[0]
        assert np.any(ri != 0)
        assert_parity(dn.inelastic(Einel[ki])[0], ri, what=f"C5 nuclide {index} ({spec[1]}) inelastic")


def test_shapes_of_the_sampled_c5_nuclides():
    assert {synth.c5_library(300)[i][1] for i in (0, 1, 2, 7, 11)} == {"light", "medium", "heavy"}


def test_freegas_column_with_non_positive_elastic_xs_is_zero(scatt, oracle):
    """Below the free-gas cutoff k_elastic leaves the column to the free-gas kernels, which only write columns whose
    interpolated elastic cross section is positive (scatt_interp_distro returns distro = ZERO otherwise,
    src/scattdata_header.F90:414-419): such a column must come back as zeros, not as whatever the buffer held."""
    nuc, e_bins, params, Ein = synth.c3_h1_freegas(n_ein=40)
    nuc.elastic = nuc.elastic.copy()
    nuc.elastic[:30] = 0.0                       # sigma_s = 0 over the low end of the grid
    dn, rn = _pair(scatt, oracle, nuc, e_bins, params)
    E = np.array([nuc.energy[3] * 1.3, nuc.energy[12], Ein[25], Ein[30]])
    for _ in range(2):                           # the second call re-uses the pool's (now dirty) memory
        got = dn.elastic(E)
    ref = rn.elastic(E)
    assert np.all(ref[:2] == 0.0) and np.any(ref[2:] != 0.0)
    assert np.array_equal(got[:2], ref[:2])
    assert_parity(got, ref, what="free gas with a vanishing elastic xs")
