"""Row N1: synthetic ACE type-1 files (ndpp_b200/acefile.py).

No Fortran compiler exists here, so the reference's reader (src/ace.F90) cannot consume the files in this
container; the tests pin the writer against the restated parse (`read_ace`, which follows read_ace_table,
read_reactions, read_angular_dist, get_energy_dist and length_energy_dist line by line), against the fixed-width
formats the reference reads them with, and against the oracle: the moments of a nuclide read back from its
file agree with those of the in-memory nuclide to the 13 digits the file carries."""
import re

import numpy as np
import pytest

from ndpp_b200 import ace, acefile, synth


def _law3(awr, Q):
    return ace.DistEnergy(law=3, data=np.array([(awr + 1.0) / awr * abs(Q), (awr / (awr + 1.0)) ** 2]))


def _law61_full(lab=False):
    """Law 61 continuum with a tabular angular table behind every outgoing energy."""
    rng = np.random.default_rng(161)
    energy = np.geomspace(1e-11, 20.0, 90)
    thr = int(np.searchsorted(energy, 2.0)) + 1
    e_in = np.array([energy[thr - 1], 6.0, 20.0])
    blocks, locs = [], []
    pos = 2 + 2 * len(e_in)
    for E in e_in:
        NP = int(rng.integers(4, 8))
        Eout = np.linspace(0.0, 0.5 * E, NP)
        pdf = np.exp(-Eout / (0.2 * E)); pdf /= np.sum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(Eout))
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(Eout))])
        ang, LC = [], []
        apos = pos + 2 + 4 * NP
        for j in range(NP):
            npa = int(rng.integers(3, 9))
            mu = np.linspace(-1, 1, npa)
            p = np.exp(rng.uniform(0, 2) * mu); p /= np.sum(0.5 * (p[1:] + p[:-1]) * np.diff(mu))
            c = np.concatenate([[0.0], np.cumsum(0.5 * (p[1:] + p[:-1]) * np.diff(mu))])
            LC.append(float(apos))
            blk = np.concatenate([[float(1 + (j % 2)), float(npa)], mu, p, c])
            ang.append(blk); apos += len(blk)
        locs.append(pos)
        blocks.append(np.concatenate([[2.0, float(NP)], Eout, pdf, cdf, LC] + ang))
        pos = apos
    data = np.concatenate([[0.0, float(len(e_in))], e_in, np.asarray(locs, float)] + blocks)
    sig = np.linspace(0.1, 1.5, len(energy) - thr + 1)
    pv = ace.Tab1(x=np.array([e_in[0], 20.0]), y=np.array([0.8, 1.0]))
    r61 = ace.Reaction(MT=91, Q_value=-1.9, threshold=thr, scatter_in_cm=not lab, sigma=sig,
                       edist=ace.DistEnergy(law=61, data=data, p_valid=pv))
    # law 9 with a nested law 66 behind it, an energy-dependent yield, and a lab 32-equiprobable adist
    d9 = np.concatenate([[0.0, 3.0], e_in, [0.3, 0.6, 1.1], [0.4]])
    e9 = ace.DistEnergy(law=9, data=d9, p_valid=ace.Tab1(x=np.array([e_in[0], 20.0]), y=np.array([0.5, 0.5])),
                        next=ace.DistEnergy(law=66, data=np.array([3.0, 4.5]),
                                            p_valid=ace.Tab1(x=np.array([e_in[0], 20.0]), y=np.array([0.5, 0.5]))))
    r9 = ace.Reaction(MT=16, Q_value=-0.9, threshold=thr, scatter_in_cm=False, multiplicity=2, sigma=sig,
                      multiplicity_E=ace.Tab1(x=np.array([e_in[0], 20.0]), y=np.array([2.0, 2.4]),
                                              nbt=np.array([2], np.int32), int=np.array([2], np.int32)),
                      adist=synth.make_adist([e_in[0], 20.0], [ace.ANGLE_32_EQUI, ace.ANGLE_TABULAR], [0.5, 1.0], NP_tab=7),
                      edist=e9)
    r51 = ace.Reaction(MT=51, Q_value=-0.3, threshold=thr - 3, scatter_in_cm=True,
                       sigma=np.full(len(energy) - thr + 4, 0.2),
                       adist=synth.make_adist([energy[thr - 4], 9.0, 20.0], [ace.ANGLE_ISOTROPIC, ace.ANGLE_TABULAR, ace.ANGLE_32_EQUI],
                                              [0.0, 0.7, 1.2], NP_tab=11),
                       edist=_law3(55.3, -0.3))
    capture = ace.Reaction(MT=102, Q_value=6.1, threshold=1, multiplicity=0, scatter_in_cm=False, sigma=np.full(len(energy), 0.01))
    el = ace.Reaction(MT=2, threshold=1,
                      adist=synth.make_adist([1e-11, 1.0, 20.0], [ace.ANGLE_ISOTROPIC, ace.ANGLE_32_EQUI, ace.ANGLE_TABULAR],
                                             [0.0, 0.4, 2.0], NP_tab=9))
    return ace.Nuclide(awr=55.3, kT=2.5301e-8, energy=energy, elastic=np.full(len(energy), 3.0),
                       reactions=[el, r61, r9, r51, capture], name="26056.70c")


def _same_tab1(a, b, rtol):
    if a is None or b is None:
        return a is None and b is None
    return (np.allclose(a.x, b.x, rtol=rtol, atol=0) and np.allclose(a.y, b.y, rtol=rtol, atol=0)
            and list(np.asarray(a.nbt).ravel()) == list(np.asarray(b.nbt).ravel())
            and list(np.asarray(a.int).ravel()) == list(np.asarray(b.int).ravel()))


def _assert_same_nuclide(a, b, rtol=6e-13):          # 1PE20.12 = 13 significant digits
    assert abs(a.awr - b.awr) <= 1e-6 * a.awr and abs(a.kT - b.kT) <= 1e-4 * a.kT + 1e-30   # header: F12.6, E12.4
    assert np.allclose(a.energy, b.energy, rtol=rtol, atol=0) and np.allclose(a.elastic, b.elastic, rtol=rtol, atol=0)
    assert len(a.reactions) == len(b.reactions)
    for ra, rb in zip(a.reactions, b.reactions):
        assert (ra.MT, ra.threshold, bool(ra.scatter_in_cm)) == (rb.MT, rb.threshold, bool(rb.scatter_in_cm))
        if ra.multiplicity_E is None:
            assert rb.multiplicity_E is None and ra.multiplicity == rb.multiplicity
        else:
            assert rb.multiplicity > 100 and _same_tab1(ra.multiplicity_E, rb.multiplicity_E, rtol)
        assert np.isclose(ra.Q_value, rb.Q_value, rtol=rtol, atol=0)
        if ra.MT != ace.ELASTIC:
            assert np.allclose(ra.sigma, rb.sigma, rtol=rtol, atol=0)
        assert (ra.adist is None) == (rb.adist is None)
        if ra.adist is not None:
            pa = acefile.pack_adist(ra.adist)
            assert np.allclose(pa.energy, rb.adist.energy, rtol=rtol, atol=0)
            assert list(pa.type) == list(rb.adist.type) and list(pa.location) == list(rb.adist.location)
            assert np.allclose(pa.data, rb.adist.data, rtol=rtol, atol=0)
        ea, eb = ra.edist, rb.edist
        while ea is not None:
            assert eb is not None and ea.law == eb.law
            assert np.allclose(ea.data, eb.data, rtol=rtol, atol=0) and len(ea.data) == len(eb.data)
            if ea.p_valid is not None:
                assert _same_tab1(ea.p_valid, eb.p_valid, rtol)
            ea, eb = ea.next, eb.next
        assert eb is None


def _nuclides():
    heavy = synth.heavy_nuclide(n_grid=600, n_levels=6, seed=5)
    heavy = heavy[0] if isinstance(heavy, tuple) else heavy
    h1 = synth.c3_h1_freegas()[0]
    return {"heavy": heavy, "h1": h1, "law61_cm": _law61_full(False), "law61_lab": _law61_full(True)}


@pytest.mark.parametrize("which", ["heavy", "h1", "law61_cm", "law61_lab"])
def test_write_read_round_trip(tmp_path, which):
    nuc = _nuclides()[which]
    p1, p2 = str(tmp_path / "a.ace"), str(tmp_path / "b.ace")
    acefile.write_ace(nuc, p1, zaid=26056, name="26056.70c")
    back = acefile.read_ace(p1)
    _assert_same_nuclide(nuc, back)
    acefile.write_ace(back, p2, zaid=26056, name="26056.70c")
    assert open(p1).read() == open(p2).read()          # read -> write is the identity on a file
    _assert_same_nuclide(back, acefile.read_ace(p2), rtol=0.0)


def test_fixed_width_layout_of_the_reference_formats(tmp_path):
    """Header (A10,2G12.0,1X,A10), A70,A10, 4 x 4(I7,F11.0), 6 x 8I9, XSS 4G20.0 (src/ace.F90:288-308)."""
    nuc = _nuclides()["law61_cm"]
    p = str(tmp_path / "t.ace")
    listing = acefile.write_ace(nuc, p, zaid=26056, name="26056.70c")
    lines = open(p).read().split("\n")
    assert lines[0][:10] == "26056.70c " and len(lines[0]) == 45 and lines[0][34] == " "
    assert float(lines[0][10:22]) == pytest.approx(nuc.awr, rel=1e-6) and float(lines[0][22:34]) == pytest.approx(nuc.kT, rel=1e-4)
    assert len(lines[1]) == 80
    for ln in lines[2:6]:
        assert len(ln) == 4 * 18
    for ln in lines[6:12]:
        assert len(ln) == 72 and re.fullmatch(r"( *\d+){8}", ln)
    nxs = [int(lines[6 + r][9 * k:9 * k + 9]) for r in range(2) for k in range(8)]
    jxs = [int(lines[8 + r][9 * k:9 * k + 9]) for r in range(4) for k in range(8)]
    body = [ln for ln in lines[12:] if ln]
    assert all(len(ln) % 20 == 0 and len(ln) <= 80 for ln in body)
    assert sum(len(ln) // 20 for ln in body) == nxs[0] == jxs[21]
    assert nxs[1] == 26056 and nxs[2] == len(nuc.energy) and nxs[3] == len(nuc.reactions) - 1 and nxs[4] == 3
    assert jxs[0] == 1 and all(jxs[k] > jxs[k - 1] for k in range(3, 11))
    assert listing["name"] == "26056.70c" and listing["location"] == 1


def test_locators_are_dlw_relative(tmp_path):
    """The row locators in the file are relative to JXS(11) (L_file = L_data + LOCC + lid) and are re-based by
    the reader exactly as length_energy_dist does (src/ace.F90:1131-1151)."""
    nuc = _nuclides()["heavy"]
    nxs, jxs, xss = acefile.build_xss(nuc)
    cont = [r for r in nuc.reactions if r.edist is not None and r.edist.law == 44][0]
    i = [r for r in nuc.reactions[1:] if r.edist is not None].index(cont)
    LOCC = int(xss[jxs[9] - 1 + i])
    LDIS = jxs[10]
    assert int(xss[LDIS + LOCC - 1]) == 44                      # LAW
    IDAT = int(xss[LDIS + LOCC])
    lc = LDIS + IDAT - 2                                        # data = XSS(lc+1:)
    d = cont.edist.data
    NE = int(d[1])
    for k in range(NE):
        L_file = int(xss[lc + 2 + NE + k])
        L_data = int(d[2 + NE + k])
        # XSS(JXS(11) + L - 1) is the row's INTT', as the ACE format defines L
        assert xss[LDIS + L_file - 2] == d[L_data] and xss[LDIS + L_file - 1] == d[L_data + 1]


def test_refusals():
    nuc, _, _ = synth.c1_fixture()
    with pytest.raises(ValueError, match="angular distribution without an energy law"):
        acefile.build_xss(nuc)                                   # the unit-test fixture is not an ACE-shaped nuclide
    from tests.test_gpu_parity import _law61_nuclide
    with pytest.raises(ValueError, match="one table per outgoing energy"):
        acefile.build_xss(_law61_nuclide(False))


def test_cross_sections_xml(tmp_path):
    import xml.etree.ElementTree as ET
    nuc = _nuclides()["h1"]
    t = acefile.write_ace(nuc, str(tmp_path / "h1.ace"), zaid=1001, name="1001.70c")
    acefile.write_cross_sections_xml([t], str(tmp_path / "cross_sections.xml"))
    root = ET.parse(str(tmp_path / "cross_sections.xml")).getroot()
    assert root.tag == "cross_sections" and root.find("filetype").text == "ascii"
    e = root.find("ace_table")
    assert e.get("name") == "1001.70c" and int(e.get("zaid")) == 1001 and int(e.get("location")) == 1
    assert float(e.get("awr")) == pytest.approx(nuc.awr, rel=1e-6)


@pytest.mark.parametrize("which", ["heavy", "law61_cm"])
def test_moments_of_the_file_match_the_nuclide(tmp_path, which):
    """The oracle on the nuclide parsed from the file = the oracle on the in-memory nuclide, to the file's digits."""
    from oracle import pyoracle
    nuc = _nuclides()[which]
    p = str(tmp_path / "n.ace")
    acefile.write_ace(nuc, p, name="26056.70c")
    back = acefile.read_ace(p)
    back.freegas_cutoff = nuc.freegas_cutoff = 0.0
    e_bins = synth.group_structure(12, 1e-6, 20.0)
    params = ace.Params(order=3, mu_bins=201)
    Ein = np.array([1e-4, 0.5, 2.5, 7.0, 19.0])
    outs = []
    for n in (nuc, back):
        rn = pyoracle.RefNuclide(n, e_bins, params)
        rn.convert_distro()
        outs.append((rn.elastic(Ein), rn.inelastic(Ein)[0]))
    for a, b in zip(*outs):
        assert np.any(a != 0)
        assert np.allclose(a, b, rtol=1e-8, atol=1e-11)
