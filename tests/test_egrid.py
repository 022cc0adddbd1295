"""The incoming-energy grid builders: create_Ein_grid (src/scatt.F90:166-536) and sab_egrid
(src/sab.F90:460-568).

Parity unpinned -- the reference holds no test for these routines.  Three statements of the same Fortran text are held
against each other, bit for bit:
  * oracle/egrid_ref.c: the literal chain of two-pointer merges (src/array_merge.F90), log / exp of the C library;
  * ndpp_b200/egrid.py: written first and separately, every merge as a sorted union (numpy), math.log / math.exp;
  * the device path (csrc/kernels_egrid.cuh): one thread per candidate point, radix sort, compaction, lm::log_ / lm::exp_.
"""
import os

import numpy as np
import pytest

from ndpp_b200 import ace, egrid, synth
from oracle import pyoracle


def _nuclides():
    eb = synth.group_structure(70)
    p7 = ace.Params(order=7, mu_bins=2001, nuscatter=False)
    yield "heavy, 4 levels", synth.heavy_nuclide(n_grid=300, n_levels=4, seed=7, n_ein_cont=6, np_cont=10, n_el_adist=8,
                                                 n_lvl_adist=4, np_lvl=9), eb, p7
    yield "heavy, 12 levels", synth.heavy_nuclide(n_grid=2000, n_levels=12, seed=11, n_ein_cont=8, np_cont=12, n_el_adist=10,
                                                  n_lvl_adist=5, np_lvl=9), eb, p7
    nuc, eb3, p3, _ = synth.c3_h1_freegas()
    yield "H-1 free gas (elastic only)", nuc, eb3, p3
    nuc, eb1, p1 = synth.c1_fixture()
    yield "C1 fixture", nuc, eb1, p1
    # group structure that stops below the top of the nuclide grid and of the channels' grids: the cuts at E_bins(size)
    # (every group structure starts at zero, src/ndpp.F90:229: the zero -> MIN_EIN rule of merge is in every case)
    eb0 = synth.group_structure(12, 1.0e-7, 3.0)
    yield "12 groups, top 3 MeV", synth.heavy_nuclide(n_grid=500, n_levels=3, seed=5, n_ein_cont=6, np_cont=10,
                                                              n_el_adist=8, n_lvl_adist=4, np_lvl=9), eb0, p7


NUCLIDES = list(_nuclides())


def _same(a, b, what):
    assert (a is None) == (b is None), what
    if a is None:
        return
    assert len(a) == len(b), f"{what}: {len(a)} vs {len(b)} points"
    assert np.array_equal(a.view(np.uint64), b.view(np.uint64)), \
        f"{what}: {np.count_nonzero(a != b)} of {len(a)} points differ, max rel {np.max(np.abs(a - b) / np.abs(b)):.3e}"


@pytest.mark.parametrize("case", range(len(NUCLIDES)), ids=[c[0] for c in NUCLIDES])
def test_oracle_chain_of_merges_equals_the_sorted_union(case):
    name, nuc, eb, params = NUCLIDES[case]
    rn = pyoracle.RefNuclide(nuc, eb, params)
    try:
        el, inel = rn.create_ein_grid()
    finally:
        rn.close()
    el2, inel2 = egrid.create_Ein_grid(nuc, eb)
    _same(el, el2, name + " Ein_el")
    _same(inel, inel2, name + " Ein_inel")
    assert np.all(np.diff(el) > 0) and (inel is None or np.all(np.diff(inel) > 0))
    # add_one_more_point: the single-precision literal 1.0E-3 promoted to double (:438)
    assert el[-1] == el[-2] * (1.0 + float(np.float32(1.0e-3)))


def test_oracle_other_extension_counts():
    name, nuc, eb, params = NUCLIDES[0]
    rn = pyoracle.RefNuclide(nuc, eb, params)
    try:
        el, inel = rn.create_ein_grid(extend_pts=7, inel_extend_pts=4)
    finally:
        rn.close()
    el2, inel2 = egrid.create_Ein_grid(nuc, eb, extend_pts=7, inel_extend_pts=4)
    _same(el, el2, "Ein_el")
    _same(inel, inel2, "Ein_inel")


GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "egrid_vectors.npz")


def _golden_inputs():
    eb = synth.group_structure(70)
    nuc, _, p7 = NUCLIDES[0][1], None, NUCLIDES[0][3]
    nuc3, eb3, p3, _ = synth.c3_h1_freegas()
    return eb, nuc, p7, nuc3, eb3, p3


def test_oracle_reproduces_the_committed_grids():
    """tests/golden/egrid_vectors.npz (scripts/make_egrid_golden.py: the numpy statement, committed) against the oracle's
    literal chain of merges."""
    g = np.load(GOLDEN)
    eb, nuc, p7, nuc3, eb3, p3 = _golden_inputs()
    rn = pyoracle.RefNuclide(nuc, eb, p7)
    try:
        el, inel = rn.create_ein_grid()
        _same(el, g["heavy4_el"], "Ein_el"); _same(inel, g["heavy4_inel"], "Ein_inel")
        el, inel = rn.create_ein_grid(extend_pts=7, inel_extend_pts=4)
        _same(el, g["heavy4_el_7_4"], "Ein_el 7/4"); _same(inel, g["heavy4_inel_7_4"], "Ein_inel 7/4")
    finally:
        rn.close()
    rn = pyoracle.RefNuclide(nuc3, eb3, p3)
    try:
        _same(rn.create_ein_grid()[0], g["h1_el"], "H-1 Ein_el")
    finally:
        rn.close()
    _same(pyoracle.sab_egrid(synth.c4_sab(mode="skewed", elastic="coherent", n_ein=20, n_eout=12), eb), g["sab_skewed_coherent"], "sab skewed")
    _same(pyoracle.sab_egrid(synth.c4_sab(mode="cont"), eb, sab_epts_per_bin=0), g["sab_cont_0"], "sab cont")


@pytest.mark.gpu
def test_device_reproduces_the_committed_grids():
    """The CUDA path against the committed vectors directly (not through the restatement)."""
    from ndpp_b200 import scatt
    g = np.load(GOLDEN)
    eb, nuc, p7, nuc3, eb3, p3 = _golden_inputs()
    dn = scatt.DeviceNuclide(nuc, eb, p7)
    try:
        el, inel, _ = dn.create_ein_grid()
        _same(el, g["heavy4_el"], "Ein_el"); _same(inel, g["heavy4_inel"], "Ein_inel")
        el, inel, _ = dn.create_ein_grid(extend_pts=7, inel_extend_pts=4)
        _same(el, g["heavy4_el_7_4"], "Ein_el 7/4"); _same(inel, g["heavy4_inel_7_4"], "Ein_inel 7/4")
    finally:
        dn.clear()
    dn = scatt.DeviceNuclide(nuc3, eb3, p3)
    try:
        _same(dn.create_ein_grid()[0], g["h1_el"], "H-1 Ein_el")
    finally:
        dn.clear()
    for key, sab, epts in (("sab_skewed_coherent", synth.c4_sab(mode="skewed", elastic="coherent", n_ein=20, n_eout=12), 10),
                           ("sab_cont_0", synth.c4_sab(mode="cont"), 0)):
        ds = scatt.DeviceSab(sab)
        try:
            _same(ds.egrid(eb, sab_epts_per_bin=epts)[0], g[key], key)
        finally:
            ds.clear()


SAB_CASES = [(m, e) for m in ("skewed", "equal", "cont") for e in (None, "coherent", "incoherent")]


@pytest.mark.parametrize("mode,elastic", SAB_CASES)
def test_oracle_sab_egrid(mode, elastic):
    sab = synth.c4_sab(mode=mode, elastic=elastic)
    eb = synth.group_structure(70)
    for epts in (10, 0):
        _same(pyoracle.sab_egrid(sab, eb, sab_epts_per_bin=epts), egrid.sab_egrid(sab, eb, sab_epts_per_bin=epts),
              f"sab_egrid {mode} {elastic} SAB_EPTS_PER_BIN={epts}")


# ---- device ------------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("case", range(len(NUCLIDES)), ids=[c[0] for c in NUCLIDES])
def test_device_create_ein_grid_equals_the_oracle(case):
    from ndpp_b200 import scatt
    name, nuc, eb, params = NUCLIDES[case]
    rn = pyoracle.RefNuclide(nuc, eb, params)
    dn = scatt.DeviceNuclide(nuc, eb, params)
    try:
        el, inel = rn.create_ein_grid()
        d_el, d_inel, status = dn.create_ein_grid()
        assert status == 0
        _same(d_el, el, name + " Ein_el")
        _same(d_inel, inel, name + " Ein_inel")
        d_el, d_inel, status = dn.create_ein_grid(extend_pts=7, inel_extend_pts=4)
        el, inel = rn.create_ein_grid(extend_pts=7, inel_extend_pts=4)
        _same(d_el, el, name + " Ein_el (7 / 4 points)")
        _same(d_inel, inel, name + " Ein_inel (7 / 4 points)")
    finally:
        dn.clear()
        rn.close()


@pytest.mark.gpu
def test_device_grid_feeds_the_integrators_without_a_host_copy():
    """The device-resident grids go straight into elastic_dev / inelastic_dev; the moments equal those of the host-grid
    call bit for bit."""
    import ctypes as C
    import torch
    from ndpp_b200 import scatt
    name, nuc, eb, params = NUCLIDES[0]
    dn = scatt.DeviceNuclide(nuc, eb, params)
    try:
        el, inel, _ = dn.create_ein_grid()
        (p_el, n_el), (p_in, n_in), _ = dn.create_ein_grid(host=False)
        assert (n_el, n_in) == (len(el), len(inel))
        G, L = len(eb) - 1, params.order + 1
        out = torch.empty((n_in, G, L), dtype=torch.float64, device="cuda")
        scatt.check(dn.lib.ndppgpu_inelastic_dev(dn.h, C.c_void_p(p_in), n_in, C.c_void_p(out.data_ptr()), None), dn.ctx.h)
        torch.cuda.synchronize()
        ref, _ = dn.inelastic(inel)
        assert np.array_equal(out.cpu().numpy(), ref)
    finally:
        dn.clear()


@pytest.mark.gpu
def test_device_c2_grid_full_size():
    """The C2 nuclide (20 000 grid points, 40 levels): 24 238 elastic and 87 029 inelastic incoming energies."""
    from ndpp_b200 import scatt
    nuc, eb, params = synth.c2_u238()[:3]
    rn = pyoracle.RefNuclide(nuc, eb, params)
    dn = scatt.DeviceNuclide(nuc, eb, params)
    try:
        el, inel = rn.create_ein_grid()
        d_el, d_inel, status = dn.create_ein_grid()
        assert status == 0
        _same(d_el, el, "C2 Ein_el")
        _same(d_inel, inel, "C2 Ein_inel")
    finally:
        dn.clear()
        rn.close()


@pytest.mark.gpu
@pytest.mark.parametrize("mode,elastic", SAB_CASES)
def test_device_sab_egrid_equals_the_oracle(mode, elastic):
    from ndpp_b200 import scatt
    sab = synth.c4_sab(mode=mode, elastic=elastic)
    eb = synth.group_structure(70)
    ds = scatt.DeviceSab(sab)
    try:
        for epts in (10, 0):
            got, status = ds.egrid(eb, sab_epts_per_bin=epts)
            assert status == 0
            _same(got, pyoracle.sab_egrid(sab, eb, sab_epts_per_bin=epts), f"sab_egrid {mode} {elastic} SAB_EPTS_PER_BIN={epts}")
    finally:
        ds.clear()


@pytest.mark.gpu
def test_device_grid_status_word():
    """The status word.  A NaN critical energy of add_inelastic_Eins (a reaction with positive Q at group edges far below
    |Q|: negative discriminant, src/scatt.F90:489) leaves its points out and sets bit 1; the reference's merge drops an
    all-NaN array too (every comparison fails, array_merge.F90:78-100), so device, numpy restatement and the oracle's
    literal chain agree in every bit.  A value repeated inside an input array is dropped and sets bit 2."""
    import copy
    from ndpp_b200 import scatt
    name, nuc, eb, params = NUCLIDES[0]
    nuc2 = copy.deepcopy(nuc)
    lvl = [r for r in nuc2.reactions if 51 <= r.MT <= 90][0]
    lvl.Q_value = 0.75
    dn = scatt.DeviceNuclide(nuc2, eb, params)
    try:
        d_el, d_inel, status = dn.create_ein_grid()
    finally:
        dn.clear()
    assert status & 2
    el, inel = egrid.create_Ein_grid(nuc2, eb)
    _same(d_el, el, "positive Q: Ein_el")
    _same(d_inel, inel, "positive Q: Ein_inel")
    assert np.all(np.isfinite(d_inel))
    rn = pyoracle.RefNuclide(nuc2, eb, params)
    try:
        r_el, r_inel = rn.create_ein_grid()
    finally:
        rn.close()
    _same(d_el, r_el, "positive Q: Ein_el vs the oracle")
    _same(d_inel, r_inel, "positive Q: Ein_inel vs the oracle")
    eb2 = np.concatenate([eb[:5], eb[4:]])          # one edge twice
    dn = scatt.DeviceNuclide(nuc, eb2, params)
    try:
        d_el, d_inel, status = dn.create_ein_grid()
    finally:
        dn.clear()
    assert status & 4
    assert np.all(np.diff(d_el) > 0) and np.all(np.diff(d_inel) > 0)
