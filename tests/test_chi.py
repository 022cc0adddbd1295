"""Row N4 (chi): fission-spectrum integration, src/chi.F90 + src/chidata_header.F90.

The reference holds no test for these routines (parity unpinned): the oracle restatement (oracle/chi_ref.c) is
pinned here by analytic properties of each law, and the CUDA path (`ndppgpu_chi` through ndpp_b200.chi.calc_chi)
is compared with it.  Tolerance: 1e-9 relative or 1e-12 absolute (BASELINE.json); the laws go through
exp / erf / sinh, where libdevice and glibc may differ in the last bits, so bit equality is not asked for.
"""
import copy
import math

import numpy as np
import pytest

from ndpp_b200 import ace, synth
from ndpp_b200 import chi as hostchi
from tests.util import assert_parity


@pytest.fixture(scope="module")
def oracle():
    from oracle import pyoracle
    pyoracle.lib()
    return pyoracle


E_BINS = synth.group_structure(70)
# general evaporation spectrum (law 5, "Not Yet Supported"): [TAB1 of T(E), NET, X(NET)]
_LAW5 = np.concatenate([synth._tab1_data([1e-11, 20.0], [1.0, 1.0]), [3.0, 0.1, 0.5, 0.9]])


def _single(law, data, threshold=1, n_grid=50, mt=18, nxt=None, p_valid=None):
    energy = np.geomspace(1e-11, 20.0, n_grid)
    energy[0], energy[-1] = 1e-11, 20.0
    ed = ace.DistEnergy(law=law, data=np.asarray(data, float), p_valid=p_valid, next=nxt)
    rx = ace.Reaction(MT=mt, threshold=threshold, sigma=np.full(n_grid - threshold + 1, 2.0), scatter_in_cm=False, edist=ed)
    return ace.Nuclide(awr=235.0, kT=2.53e-8, energy=energy, elastic=np.ones(n_grid),
                       reactions=[ace.Reaction(MT=2, threshold=1), rx], nu_t_type=1, nu_t_data=np.array([2.0, 2.4, 0.1]))


def test_law4_is_the_difference_of_the_cdf(oracle):
    e_in = np.array([1e-11, 1.0, 20.0])
    edges = np.array([0.0, 1e-3, 0.1, 0.5, 1.0, 2.0, 5.0, 20.0])
    rows = []
    for k in range(3):
        pdf = np.exp(-edges / (1.0 + k))
        cdf = synth._lin_cdf(edges, pdf)
        rows.append((2, edges, pdf / cdf[-1], cdf / cdf[-1], np.zeros(0), np.zeros(0)))
    nuc = _single(4, synth.make_law44(e_in, rows))
    E, t, p, d = oracle.calc_chi(nuc, edges)          # groups = the table's own E_out intervals
    assert np.array_equal(E, e_in) and d.shape[0] == 0
    for k in range(3):
        assert np.allclose(p[k], np.diff(rows[k][3]), rtol=0, atol=2e-16)
    assert np.allclose(p.sum(axis=1), 1.0, atol=1e-15) and np.allclose(t, p, atol=1e-15)
    # nearest-row rule (x > 0.5 picks the upper row), :287-291
    E2, _, p2, _ = oracle.calc_chi(nuc, edges, np.array([0.4, 0.6, 10.4, 10.6]))
    assert np.array_equal(p2[0], p[0]) and np.array_equal(p2[1], p[1])
    assert np.array_equal(p2[2], p[1]) and np.array_equal(p2[3], p[2])


def _maxwell_cum(E, T):
    return T ** 1.5 * (0.5 * math.sqrt(math.pi) * math.erf(math.sqrt(E / T)) - math.sqrt(E / T) * math.exp(-E / T))


def test_law7_matches_the_analytic_maxwell_integral(oracle):
    T, U = 1.3, -30.0
    nuc = _single(7, np.concatenate([synth._tab1_data([1e-11, 20.0], [T, T]), [U]]))
    Ein = np.array([1e-6, 2.0, 14.0])
    _, _, p, _ = oracle.calc_chi(nuc, E_BINS, Ein)
    cum = np.array([_maxwell_cum(e, T) for e in E_BINS])
    ref = np.diff(cum) / (cum[-1] - cum[0])
    for k in range(len(Ein)):
        # the reference subtracts two O(1) primitives, so its small groups carry ~1e-16 absolute round-off;
        # it also uses PI = 3.1415926535898 (1e-14 relative)
        assert np.allclose(p[k], ref, rtol=1e-12, atol=1e-15)


def test_law9_matches_the_analytic_evaporation_integral(oracle):
    T, U = 0.9, 1.5
    nuc = _single(9, np.concatenate([synth._tab1_data([1e-11, 20.0], [T, T]), [U]]))
    Ein = np.array([1.0, 1.5, 4.0, 20.0])
    _, _, p, _ = oracle.calc_chi(nuc, E_BINS, Ein)
    assert np.all(p[:2] == 0.0)                       # Ein <= U: the function returns before normalising (:387)
    for k in (2, 3):
        top = Ein[k] - U
        e = np.minimum(E_BINS, top)
        cum = -T * (e + T) * np.exp(-e / T)
        ref = np.diff(cum) / (cum[-1] - cum[0])
        assert np.allclose(p[k], ref, rtol=1e-11, atol=1e-15)
        assert np.all(p[k][E_BINS[:-1] >= top] == 0.0)


def test_reference_quirks_are_reproduced(oracle):
    # a law that only warns leaves zeros, and the final normalisation turns them into 0 * (1/0) = NaN (:483-492)
    nuc = _single(5, _LAW5)
    _, t, p, _ = oracle.calc_chi(nuc, E_BINS, np.array([1.0, 2.0]))
    assert np.all(np.isnan(p)) and np.all(np.isnan(t))
    # law 7 replaces a group edge above Ein - U by U itself (:358,362): with U < 0 that is sqrt(negative)
    nuc = _single(7, np.concatenate([synth._tab1_data([1e-11, 20.0], [1.3, 1.3]), [-5.0]]))
    _, _, p, _ = oracle.calc_chi(nuc, E_BINS, np.array([1.0]))
    assert np.all(np.isnan(p))
    # below the reaction threshold prob = 0 (:199-201) and chi_total is 0 -> left un-normalised (norm > 0 fails)
    nuc = _single(9, np.concatenate([synth._tab1_data([1e-11, 20.0], [0.9, 0.9]), [0.0]]), threshold=45, mt=19)
    _, t, p, _ = oracle.calc_chi(nuc, E_BINS, np.array([nuc.energy[40], nuc.energy[47]]))
    assert np.all(t[0] == 0.0) and p[0].sum() == 0.0 and abs(t[1].sum() - 1.0) < 1e-14


def test_chi_total_combination_rule(oracle):
    """chi_total = chi_prompt (1 + prob_last) + beta sum_k yield_k chi_delay_k, then normalised (src/chi.F90:125-147)."""
    nuc = synth.fissile_total()
    E, t, p, d = oracle.calc_chi(nuc, E_BINS)
    L = oracle.lib()
    for k in (0, 3, len(E) - 1):
        Ein = E[k]
        beta = L.ref_interpolate_tab1(oracle.dp(oracle.f64(nuc.nu_d_data)), Ein) / \
            L.ref_interpolate_tab1(oracle.dp(oracle.f64(nuc.nu_t_data)), Ein)
        comb = p[k] * (1.0 + 1.0)                     # one MT 18 reaction: prob = fission / fission = 1
        off = 0
        for g in range(nuc.n_precursor):
            blk = oracle.f64(nuc.nu_d_precursor_data[off + 1:])
            y = L.ref_interpolate_tab1(oracle.dp(blk), Ein)
            comb = comb + y * beta * d[g, k]
            NR = int(blk[0]); NE = int(blk[1 + 2 * NR])
            off += 1 + 2 + 2 * NR + 2 * NE
        assert np.allclose(t[k], comb / comb.sum(), rtol=1e-13, atol=1e-16)
    assert np.allclose(d.sum(axis=2), 1.0, atol=1e-14)


def test_merged_grid_and_slot_list():
    nuc = synth.fissile_partial()
    slots, pool = hostchi.chi_data(nuc)
    assert [s["law"] for s in slots] == [4, 7, 9, 11, 7] and [s["use_pvalid"] for s in slots] == [1, 0, 0, 0, 0]
    E = hostchi.chi_grid(slots)
    assert np.array_equal(E, np.array([1e-11, 0.5, 1.0, 6.0, 20.0]))
    nuc = synth.fissile_total()
    slots, pool = hostchi.chi_data(nuc)
    assert [s["delayed"] for s in slots] == [0] + [1] * 6 and [s["precursor"] for s in slots[1:]] == [1, 2, 3, 4, 5, 6]
    E = hostchi.chi_grid(slots)
    assert np.all(np.diff(E) > 0) and len(E) == 15
    assert pool[slots[0]["sigma_off"]:slots[0]["sigma_off"] + slots[0]["n_sigma"]].tolist() == hostchi.fission_xs(nuc).tolist()


# ---------------------------------------------------------------------------------------------------------------
# CUDA path against the oracle
# ---------------------------------------------------------------------------------------------------------------
def _same(got, ref, what):
    assert np.array_equal(np.isnan(got), np.isnan(ref)), what
    m = ~np.isnan(ref)
    assert_parity(got[m], ref[m], what=what)


@pytest.mark.gpu
@pytest.mark.parametrize("which", ["total", "partial"])
def test_gpu_chi_matches_oracle(oracle, which):
    nuc = synth.fissile_total() if which == "total" else synth.fissile_partial()
    for grid in (None, np.geomspace(1e-11, 20.0, 700)):
        Eg, gt, gp, gd = hostchi.calc_chi(nuc, E_BINS, E_grid=grid)
        Er, rt, rp, rd = oracle.calc_chi(nuc, E_BINS, E_grid=grid)
        assert np.array_equal(Eg, Er) and np.any(rt > 0)
        _same(gt, rt, f"chi_total {which}")
        _same(gp, rp, f"chi_prompt {which}")
        _same(gd, rd, f"chi_delay {which}")
        assert not np.isnan(rt).any()


@pytest.mark.gpu
def test_gpu_chi_quirks_and_errors(oracle):
    from ndpp_b200.capi import NdppGpuError
    nuc = _single(5, _LAW5)
    _, gt, gp, _ = hostchi.calc_chi(nuc, E_BINS, E_grid=np.array([1.0, 2.0]))
    assert np.all(np.isnan(gp)) and np.all(np.isnan(gt))
    nuc = _single(9, np.concatenate([synth._tab1_data([1e-11, 20.0], [0.9, 0.9]), [1.5]]), threshold=45, mt=19)
    grid = np.array([1.0, 1.5, nuc.energy[40], nuc.energy[47], 20.0])
    g, r = hostchi.calc_chi(nuc, E_BINS, E_grid=grid), oracle.calc_chi(nuc, E_BINS, E_grid=grid)
    for a, b in zip(g[1:3], r[1:3]):
        _same(a, b, "law 9 below U / below threshold")
    # two interpolation regions in a law-4 table: the reference's fatal_error text (:272-274)
    e_in = np.array([1e-11, 20.0])
    d = synth.make_law44(e_in, synth._watt_rows(e_in, 10))
    d = np.concatenate([[2.0, 1.0, 2.0, 2.0, 2.0], d[1:]])
    d[5 + 1 + len(e_in):5 + 1 + 2 * len(e_in)] += 4
    with pytest.raises(NdppGpuError, match="Multiple interpolation regions"):
        hostchi.calc_chi(_single(4, d), E_BINS, E_grid=np.array([1.0]))
    nuc = synth.fissile_total()
    nuc.nu_t_type = 0
    with pytest.raises(ValueError, match="No neutron emission data"):
        hostchi.calc_chi(nuc, E_BINS)


# ---------------------------------------------------------------------------------------------------------------
# committed fixtures (tests/golden/chi_vectors.npz, written by scripts/make_golden.py)
# ---------------------------------------------------------------------------------------------------------------
def _chi_gold():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "chi_vectors.npz"))


def test_chi_oracle_reproduces_committed_vectors(oracle):
    v = _chi_gold()
    for name, mk in (("total", synth.fissile_total), ("partial", synth.fissile_partial)):
        for gname in ("merged", "dense"):
            E, t, p, d = oracle.calc_chi(mk(), E_BINS, E_grid=v[f"{name}_{gname}_E"])
            # exp / erf of the host libm sit between input and output: last-bit differences between machines allowed
            assert np.allclose(t, v[f"{name}_{gname}_total"], rtol=1e-13, atol=1e-16, equal_nan=True)
            assert np.allclose(p, v[f"{name}_{gname}_prompt"], rtol=1e-13, atol=1e-16, equal_nan=True)
            assert np.allclose(d, v[f"{name}_{gname}_delay"], rtol=1e-13, atol=1e-16, equal_nan=True)


@pytest.mark.gpu
def test_gpu_chi_against_committed_vectors():
    v = _chi_gold()
    for name, mk in (("total", synth.fissile_total), ("partial", synth.fissile_partial)):
        for gname in ("merged", "dense"):
            E, t, p, d = hostchi.calc_chi(mk(), E_BINS, E_grid=v[f"{name}_{gname}_E"])
            _same(t, v[f"{name}_{gname}_total"], f"golden chi_total {name} {gname}")
            _same(p, v[f"{name}_{gname}_prompt"], f"golden chi_prompt {name} {gname}")
            _same(d, v[f"{name}_{gname}_delay"], f"golden chi_delay {name} {gname}")
