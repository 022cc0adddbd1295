"""N4: the library writer (ndpp_b200/output.py) against the reference's own reader.

The reference ships a Python reader of its binary libraries (src/utils/ndpp_data.py).  In the build
container it is imported from /root/reference (Python 2 source: `xrange` / `long` are provided) and reads
the files written here; on the GPU box, where /root/reference does not exist, the same files are checked
through this repo's restatement of that reader (output.read_library) and the committed digest."""
import builtins
import hashlib
import os

import numpy as np
import pytest

from ndpp_b200 import output

REF_READER = "/root/reference/src/utils/ndpp_data.py"


def _case(nus=True):
    rng = np.random.default_rng(42)
    eb = np.array([0.0, 1e-6, 1e-3, 0.5, 2.0, 20.0])
    NG, L = 5, 4
    Eel = np.geomspace(1e-9, 20.0, 23)
    Einel = np.geomspace(0.6, 20.02, 9)

    def mat(NE):
        m = rng.normal(size=(NE, NG, L)) * 0.1
        m[:, :, 0] = np.abs(m[:, :, 0])
        m[:, 0, :] = 0.0                      # leading zero group: window starts at 2
        m[1] = 0.0                            # an all-zero column: gmin = gmax = 0
        m[2, 3:, :] = 0.0                     # trailing zeros
        return m
    return eb, Eel, mat(len(Eel)), Einel, mat(len(Einel)), (mat(len(Einel)) if nus else None)


def _write(path, fmt, nus=True, inel=True):
    eb, Eel, el, Einel, inm, nu = _case(nus)
    with output.LibraryWriter(path, "92238.71c", 2.5301e-8, eb, 0, 3, nus, 2001, 0.002, fmt) as w:
        w.print_scatt(Eel, el, Einel if inel else None, inm if inel else None, nu if inel else None)
    return eb, Eel, el, (Einel if inel else None), (inm if inel else None), (nu if (inel and nus) else None)


@pytest.mark.parametrize("nus,inel", [(True, True), (False, True), (False, False)])
def test_binary_round_trip_own_reader(tmp_path, nus, inel):
    p = str(tmp_path / "lib.bin")
    eb, Eel, el, Einel, inm, nu = _write(p, output.BINARY, nus, inel)
    r = output.read_library(p)
    assert r["trailing_bytes"] == 0 and r["name"] == "92238.71c " and r["kT"] == 2.5301e-8 and r["NG"] == 5
    assert r["scatt_type"] == 0 and r["scatt_order"] == 3 and r["mu_bins"] == 2001 and r["thin_tol"] == 0.002
    assert np.array_equal(r["E_bins"], eb) and np.array_equal(r["Ein_el"], Eel)
    # the window keeps everything between the first and the last group of positive P0
    def windowed(m):
        out = np.zeros_like(m)
        for i in range(len(m)):
            g = np.nonzero(m[i, :, 0] > 0)[0]
            if len(g):
                out[i, g[0]:g[-1] + 1] = m[i, g[0]:g[-1] + 1]
        return out
    assert np.array_equal(r["elastic"], windowed(el))
    if inel:
        assert np.array_equal(r["Ein_inel"], Einel) and np.array_equal(r["inelastic"], windowed(inm))
        assert (r["nuinelastic"] is not None) == nus
        if nus:
            assert np.array_equal(r["nuinelastic"], windowed(nu))
    else:
        assert r["Ein_inel"] is None


def test_group_index_follows_the_reference_rule():
    Ein = np.array([1e-5, 1e-3, 0.1, 1.0, 10.0])
    gi = output.group_index(Ein, [0.0, 1e-5, 5e-3, 1.0, 20.0])
    assert gi.tolist() == [1, 1, 2, 4, 5]       # below grid -> 1; interior -> binary_search; last -> size(Ein)


def test_ascii_layout(tmp_path):
    p = str(tmp_path / "lib.txt")
    eb, Eel, el, Einel, inm, nu = _write(p, output.ASCII)
    lines = open(p).read().split("\n")
    assert lines[0] == " " * 10 + "92238.71c " + "  2.530100000000E-08" + f"{5:20d}"
    assert lines[1] == "".join(f"{v:20.12E}" for v in eb[:4]) and lines[2] == "".join(f"{v:20.12E}" for v in eb[4:])
    assert lines[3] == f"{0:20d}{3:20d}{1:20d}{0:20d}" and lines[4] == f"{2001:20d}" + "  2.000000000000E-03"
    assert lines[5] == f"{len(Eel):20d}"
    assert output._fortran_e(1.5e100) == "  1.500000000000+100" and output._fortran_e(-2e-120) == " -2.000000000000-120"
    # numbers parse back to 13 significant digits
    vals = np.array([float(x) for ln in lines[6:12] for x in ln.split()])[:len(Eel)]
    assert np.allclose(vals, Eel, rtol=1e-12)


def test_golden_digest(tmp_path):
    """The byte stream of the binary file is pinned (tests/golden/library_case.sha256, written by this
    test's own generator when the reference reader validated it in the build container)."""
    p = str(tmp_path / "lib.bin")
    _write(p, output.BINARY)
    digest = hashlib.sha256(open(p, "rb").read()).hexdigest()
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "library_case.sha256")
    assert open(gold).read().split()[0] == digest


@pytest.mark.skipif(not os.path.exists(REF_READER), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("nus", [True, False])
def test_reference_reader_reads_our_files(tmp_path, nus):
    src = open(REF_READER).read()
    builtins.xrange, builtins.long = range, int       # the reader is Python 2 source
    try:
        ns = {}
        exec(compile(src, REF_READER, "exec"), ns)
        p = str(tmp_path / "lib.bin")
        eb, Eel, el, Einel, inm, nu = _write(p, output.BINARY, nus)
        lib = ns["NDPP_lib"](p, "binary")
    finally:
        del builtins.xrange, builtins.long
    assert lib.NG == 5 and lib.kT == 2.5301e-8 and lib.scatt_order == 4 and lib.mu_bins == 2001 and lib.thin_tol == 0.002
    assert lib.nuinelastic_present == nus and not lib.chi_present
    assert np.array_equal(lib.E_bins, eb) and np.array_equal(lib.Ein_el, Eel) and np.array_equal(lib.Ein_inel, Einel)
    assert np.array_equal(lib.grp_index_el, output.group_index(Eel, eb))
    for name, m in (("elastic", el), ("inelastic", inm)) + ((("nuinelastic", nu),) if nus else ()):
        recs = getattr(lib, name)
        assert len(recs) == len(m)
        for iE, rec in enumerate(recs):
            g = np.nonzero(m[iE, :, 0] > 0)[0]
            if len(g) == 0:
                assert rec.gmin == -1 and rec.gmax == -1      # 0, 0 on file, minus one in the reader
            else:
                assert (rec.gmin, rec.gmax) == (g[0], g[-1])
                assert np.array_equal(rec.outgoing, m[iE, g[0]:g[-1] + 1])


def test_chi_record_round_trip(tmp_path):
    """print_chi (src/chi.F90:195-236 ASCII, :309-337 binary) behind the scattering data, chi_present = 1."""
    rng = np.random.default_rng(3)
    eb = np.array([0.0, 1e-6, 1.0, 20.0])
    NE, G, L, NEc, NP = 4, 3, 2, 5, 2
    Ein = np.geomspace(1e-9, 19.0, NE)
    el = rng.random((NE, G, L))
    Ec = np.geomspace(1e-11, 20.0, NEc)
    ct, cp, cd = rng.random((NEc, G)), rng.random((NEc, G)), rng.random((NP, NEc, G))
    pb, pa = str(tmp_path / "n.bin"), str(tmp_path / "n.txt")
    for path, fmt in ((pb, output.BINARY), (pa, output.ASCII)):
        with output.LibraryWriter(path, "92235.70c", 2.53e-8, eb, 0, L - 1, False, 201, 0.0, fmt, chi_present=True) as w:
            w.print_scatt(Ein, el)
            w.print_chi(Ec, ct, cp, cd)
    lib = output.read_library(pb)
    assert lib["chi_present"] and lib["trailing_bytes"] == 0
    assert np.array_equal(lib["Ein_chi"], Ec) and np.array_equal(lib["chi_total"], ct)
    assert np.array_equal(lib["chi_prompt"], cp) and np.array_equal(lib["chi_delay"], cd)
    lines = open(pa).read().split("\n")
    k = [i for i, ln in enumerate(lines) if ln == f"{NEc:20d}{NP:20d}"]
    assert len(k) == 1
    vals = np.array([float(ln[20 * j:20 * j + 20]) for ln in lines[k[0] + 1:] for j in range(len(ln) // 20)])
    want = np.concatenate([Ec, ct.ravel(), cp.ravel(), cd.ravel()])
    # every array starts on a new line (print_ascii_array), values carry 13 digits
    assert len(vals) == len(want) and np.allclose(vals, want, rtol=1e-12)
    with pytest.raises(ValueError):
        with output.LibraryWriter(pb, "x", 0.0, eb, 0, 1, False, 201, 0.0) as w:
            w.print_chi(Ec, ct, cp, cd)
