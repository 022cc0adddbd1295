"""Several GPUs behind the C-ABI (csrc/group.cuh): a nuclide sharded by E_in over a device group and the work-item
runner of a library must give the one-device call's matrices bit for bit -- the partition re-organises the work, not
the arithmetic.  Tests that need more than one GPU skip on a one-GPU box (run them with `gpurun --gpus 2`)."""
import numpy as np
import pytest

from ndpp_b200 import ace, synth
from tests.util import small_heavy

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


@pytest.fixture(scope="module")
def mods():
    from ndpp_b200 import group, library, scatt
    scatt.default_context()
    return group, library, scatt


def _grids(nuc, e_bins, n_el=257, n_in=131, seed=3):
    """E_in grids with two points above the top group edge (the top-of-grid rule, src/scatt.F90:669,770)."""
    rng = np.random.default_rng(seed)
    top = e_bins[-1]
    Eel = np.sort(np.concatenate([np.exp(rng.uniform(np.log(1e-11), np.log(top), n_el - 3)), [top, top * 1.001, top * 1.002]]))
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != 2)
    Ein = np.sort(np.concatenate([rng.uniform(thr, top, n_in - 3), [top, top * 1.001, top * 1.002]]))
    return Eel, Ein


def _compare_group_with_one_device(mods, g, nuc, e_bins, params, Eel, Ein):
    group, library, scatt = mods
    dn = scatt.DeviceNuclide(nuc, e_bins, params)
    ref_el = dn.elastic(Eel)
    ref_in, ref_nu = dn.inelastic(Ein) if Ein is not None else (None, None)
    dn.clear()
    gn = group.GroupNuclide(nuc, e_bins, params, g)
    assert np.array_equal(gn.elastic(Eel), ref_el)
    if Ein is not None:
        gi, gnu = gn.inelastic(Ein)
        assert np.array_equal(gi, ref_in)
        if params.nuscatter:
            assert np.array_equal(gnu, ref_nu)
    # the same in pieces, twice in a row (two result buffers in turn, gather on the side streams)
    gn.set_grids(Eel, Ein)
    what = 3 if Ein is not None else 1
    for _ in range(3):
        gn.integrate(what)
    gn.sync()
    e, i, n = gn.fetch()
    assert np.array_equal(e, ref_el)
    if Ein is not None:
        assert np.array_equal(i, ref_in)
    top = e_bins[-1]
    j = int(np.nonzero(Eel <= top)[0][-1])                                       # the copy rule was exercised
    assert j < len(Eel) - 1 and np.array_equal(ref_el[-1], ref_el[j]) and np.any(ref_el[j] != 0)
    gn.clear()


def test_group_of_one_device_equals_the_plain_call(mods):
    group, library, scatt = mods
    g = group.Group(1)
    assert (g.world, g.n_local, g.first) == (1, 1, 0)
    nuc = small_heavy(n_grid=500)
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7, mu_bins=2001, nuscatter=True)
    Eel, Ein = _grids(nuc, e_bins)
    _compare_group_with_one_device(mods, g, nuc, e_bins, params, Eel, Ein)
    g.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")
def test_nuclide_sharded_over_all_gpus_is_bit_identical(mods):
    """ndppgpu_group_*: E_in dealt cyclically over every GPU of the box, columns gathered with ncclSend / ncclRecv."""
    group, library, scatt = mods
    g = group.Group(0)
    assert g.world == g.n_local == _n_gpus()
    e_bins = synth.group_structure(70)
    nuc = small_heavy(n_grid=500)
    Eel, Ein = _grids(nuc, e_bins)
    _compare_group_with_one_device(mods, g, nuc, e_bins, ace.Params(order=7, mu_bins=2001, nuscatter=True), Eel, Ein)
    assert g.gathered_bytes() > 0
    # fewer columns than devices, and a single column
    _compare_group_with_one_device(mods, g, nuc, e_bins, ace.Params(order=5, mu_bins=501), Eel[-3:], Ein[-4:])
    # free gas (C3 shape): the heaviest, most uneven columns
    nuc, e_bins, params, E = synth.c3_h1_freegas(n_ein=1000)
    E = np.concatenate([E[::97], [e_bins[-1] * 1.001]])
    _compare_group_with_one_device(mods, g, nuc, e_bins, params, E, None)
    g.close()


def _library_case():
    specs = synth.c5_library(300, ne_hi=4000)[:7]
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, mu_bins=1001)
    shapes = [synth.c5_shape(s) for s in specs]
    parsed = {s[0]: synth.c5_nuclide(s) for s in specs}
    # two points above the top group edge at the end of every grid: with several tiles per matrix the predecessor of
    # the first of them may sit in another tile
    for i, (nuc, Eel, Einel) in parsed.items():
        top = e_bins[-1]
        Eel = np.concatenate([Eel[Eel <= top], [top * 1.001, top * 1.002]])
        if Einel is not None:
            Einel = np.concatenate([Einel[Einel <= top], [top * 1.001, top * 1.002]])
        parsed[i] = (nuc, Eel, Einel)
    return specs, shapes, parsed, e_bins, params


def _run_library_and_compare(mods, g, plan_world, remap):
    group, library, scatt = mods
    specs, shapes, parsed, e_bins, params = _library_case()
    G, L = len(e_bins) - 1, params.order + 1
    items, imb = library.plan(shapes, G, L, params.mu_bins, params.ne_per_grp, plan_world, tile_rows=300)
    for it in items:
        it["rank"] = remap(it["rank"])
    items.sort(key=lambda it: (it["rank"], it["nuclide"], it["matrix"], it["tile"]))
    library.set_rows(items, {i: (len(p[1]), 0 if p[2] is None else len(p[2])) for i, p in parsed.items()})
    assert max(it["n_tiles"] for it in items) > 1
    run = group.LibraryRun(g, G, L, False, items)
    opened = []

    def open_nuclide(i, ctx):
        opened.append(i)
        nuc, Eel, Einel = parsed[i]
        return scatt.DeviceNuclide(nuc, e_bins, params, ctx), Eel, (Einel if Einel is not None else np.zeros(0))
    rep = run.run(open_nuclide)
    assert rep["items"] == len(items) and rep["opens"] == len(opened) and rep["device_s_max"] > 0
    evals = 0
    for i, (nuc, Eel, Einel) in parsed.items():
        dn = scatt.DeviceNuclide(nuc, e_bins, params)
        ref = dn.elastic(Eel)
        got = run.fetch(i, 0, Eel, e_bins[-1])
        assert np.array_equal(got, ref), ("elastic", i)
        assert np.array_equal(got[-1], got[-3])
        evals += ref.size
        if Einel is not None:
            ref = dn.inelastic(Einel)[0]
            assert np.array_equal(run.fetch(i, 1, Einel, e_bins[-1]), ref), ("inelastic", i)
            evals += ref.size
        dn.clear()
    assert rep["moment_evals"] == evals
    run.close()


def test_library_tiles_on_one_gpu_equal_the_monolithic_calls(mods):
    """Tiling invariance: the plan of a four-device box run on one GPU (every item mapped to device 0) -- tiles +
    ndppgpu_library_fetch == the monolithic calls, bitwise, including the top-of-grid copy across a tile edge."""
    group, library, scatt = mods
    g = group.Group(1)
    _run_library_and_compare(mods, g, plan_world=4, remap=lambda r: 0)
    g.close()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs at least two GPUs")
def test_library_over_all_gpus_equals_the_monolithic_calls(mods):
    group, library, scatt = mods
    g = group.Group(0)
    _run_library_and_compare(mods, g, plan_world=g.world, remap=lambda r: r)
    assert g.gathered_bytes() > 0
    g.close()


def test_group_errors_are_loud(mods):
    group, library, scatt = mods
    from ndpp_b200.capi import NdppGpuError
    with pytest.raises(NdppGpuError, match="more devices"):
        group.Group(_n_gpus() + 1)
    g = group.Group(1)
    nuc, e_bins, params = synth.c1_fixture()
    gn = group.GroupNuclide(nuc, e_bins, params, g, convert=False)
    with pytest.raises(NdppGpuError, match="convert_distro"):
        gn.elastic(np.array([1.5]))
    gn.convert_distro()
    with pytest.raises(NdppGpuError, match="binary search"):
        gn.inelastic(np.array([2.7]))      # the reference aborts here (search.F90:36-38); the message crosses the threads
    assert np.all(np.isfinite(gn.inelastic(np.array([2.2]))[0]))
    gn.clear()
    g.close()
