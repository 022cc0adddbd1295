"""The C++ host layer above the C-ABI (include/ndpp_host.hpp, tools/ndpp_calc_scatt.cpp): calc_scatt and
calc_scattsab with the reference's argument lists, driven from a compiled program as the reference's
preprocess_ndpp drives the Fortran routines.  CPU: it builds with -Wall -Wextra, links against libndppgpu.so,
and fails loudly the way fatal_error does.  GPU: its moments equal the oracle's (BASELINE tolerance) and the
Python mirror's bit for bit (both are thin callers of the same library)."""
import os
import subprocess

import numpy as np
import pytest

from ndpp_b200 import ace, dump, synth
from ndpp_b200 import build as nbuild
from tests.util import assert_parity, small_heavy


@pytest.fixture(scope="module")
def tool():
    nbuild.build()
    return nbuild.build_tool()


def _run(tool, case, res, expect_ok=True):
    r = subprocess.run([tool, str(case), str(res)], capture_output=True, text=True, timeout=600)
    if expect_ok:
        assert r.returncode == 0, r.stderr
    return r


def test_tool_builds_and_links_the_c_abi(tool):
    assert os.access(tool, os.X_OK)
    out = subprocess.run(["ldd", tool], capture_output=True, text=True).stdout
    assert "libndppgpu.so" in out and "not found" not in out.split("libndppgpu.so")[1].splitlines()[0]


def test_fatal_error_convention(tool, tmp_path):
    # src/error.F90:79-154: " ERROR: <message>" on stderr, non-zero status
    r = _run(tool, tmp_path / "missing.case", tmp_path / "x.res", expect_ok=False)
    assert r.returncode != 0 and r.stderr.startswith(" ERROR: Cannot open case file")
    bad = tmp_path / "bad.case"
    np.array([7.0, 1.0]).astype("<f8").tofile(bad)
    r = _run(tool, bad, tmp_path / "x.res", expect_ok=False)
    assert r.returncode != 0 and r.stderr.startswith(" ERROR: ")


def test_no_cpu_fallback_without_a_gpu(tool, tmp_path):
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(13)
    dump.write_nuclide_case(tmp_path / "c1.case", nuc, e_bins, params.scatt_type, params.order, params.mu_bins, True,
                            Ein, Ein, params)
    r = _run(tool, tmp_path / "c1.case", tmp_path / "c1.res", expect_ok=False)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr and not os.path.exists(tmp_path / "c1.res")


def test_case_file_round_trip_layout(tmp_path):
    """The case file is the flat argument list: its length follows from the fields (guards the C++ reader's
    field order against the writer's)."""
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(5)
    p = tmp_path / "c1.case"
    dump.write_nuclide_case(p, nuc, e_bins, 0, 5, 3001, False, Ein, None, params)
    a = np.fromfile(p, dtype="<f8")
    assert a[0] == dump.KIND_NUCLIDE and a[1] == nuc.awr and a[4] == len(nuc.energy)
    assert a[-1] == 0.0 and np.array_equal(a[-1 - len(Ein):-1], Ein) and a[-2 - len(Ein)] == len(Ein)


def _library_case(tmp_path, nus, inel, three_digit_exponent=False):
    """C1 nuclide header + made-up matrices with leading / trailing zero groups and an all-zero column."""
    rng = np.random.default_rng(42)
    nuc, e_bins, params = synth.c1_fixture()
    G, L = len(e_bins) - 1, 6
    Eel = np.geomspace(1e-9, 2.9, 23)
    Einel = np.geomspace(2.0, 3.003, 9) if inel else None

    def mat(NE):
        m = rng.normal(size=(NE, G, L)) * 0.1
        m[:, :, 0] = np.abs(m[:, :, 0])
        m[1] = 0.0                      # an all-zero column: gmin = gmax = 0
        m[2, 1:, :] = 0.0               # trailing zero group
        m[3, 0, :] = 0.0                # leading zero group
        if three_digit_exponent:
            m[4, 0, 1] = -3.25e-123
            m[5, 1, 2] = 7.5e+104
        return m
    el = mat(len(Eel))
    inm = mat(len(Einel)) if inel else None
    nu = mat(len(Einel)) if (inel and nus) else None
    dump.write_nuclide_case(tmp_path / "lib.case", nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, nus, Eel, Einel, params)
    dump.write_result(tmp_path / "lib.res", el, inm, nu)
    return nuc, e_bins, Eel, el, Einel, inm, nu


@pytest.mark.parametrize("fmt", ["binary", "ascii"])
@pytest.mark.parametrize("nus,inel", [(True, True), (False, True), (False, False)])
def test_cpp_library_writer_matches_the_python_writer_byte_for_byte(tool, tmp_path, fmt, nus, inel):
    """include/ndpp_library.hpp (init_library + print_scatt) against ndpp_b200/output.py, which tests/test_output.py
    holds against the reference's own reader."""
    from ndpp_b200 import output
    nuc, e_bins, Eel, el, Einel, inm, nu = _library_case(tmp_path, nus, inel, three_digit_exponent=(fmt == "ascii"))
    cpp, py = tmp_path / f"cpp.{fmt}", tmp_path / f"py.{fmt}"
    cmd = [tool, str(tmp_path / "lib.case"), str(tmp_path / "lib.res"), "--library-only", "--library", str(cpp),
           "--name", "92238.71c", "--thin-tol", "0.002"] + (["--ascii"] if fmt == "ascii" else [])
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    with output.LibraryWriter(str(py), "92238.71c", nuc.kT, e_bins, 0, 5, nus, 3001, 0.002, fmt) as w:
        w.print_scatt(Eel, el, Einel, inm, nu)
    assert open(cpp, "rb").read() == open(py, "rb").read()
    if fmt == "binary":
        lib = output.read_library(str(cpp))
        assert lib["trailing_bytes"] == 0 and lib["name"] == "92238.71c " and lib["NG"] == len(e_bins) - 1
        assert np.array_equal(lib["Ein_el"], Eel) and (lib["Ein_inel"] is None) == (not inel)


def test_cpp_library_writer_rejects_a_mismatched_result(tool, tmp_path):
    _library_case(tmp_path, False, True)
    dump.write_result(tmp_path / "lib.res", np.zeros((3, 2, 6)))
    r = subprocess.run([tool, str(tmp_path / "lib.case"), str(tmp_path / "lib.res"), "--library-only", "--library",
                        str(tmp_path / "x.bin")], capture_output=True, text=True, timeout=120)
    assert r.returncode != 0 and r.stderr.startswith(" ERROR: Result file does not match")


@pytest.mark.gpu
def test_cpp_driver_writes_the_library_the_python_driver_writes(tool, tmp_path):
    """calc_scatt + apply_tol_scatt + thin_grid + init_library + print_scatt from the C++ driver (src/ndpp.F90:560-702)
    against ndpp_b200.driver.preprocess_nuclide: the same device calls, so the files must be identical."""
    from ndpp_b200 import driver, output
    nuc = small_heavy()
    nuc.name = "92238.71c"
    e_bins = synth.group_structure(70)
    params = ace.Params(order=5, mu_bins=2001, nuscatter=True)
    rng = np.random.default_rng(12)
    Ein = np.sort(np.exp(rng.uniform(np.log(1e-10), np.log(19.9), 300)))
    Ein_inel = Ein[Ein >= 0.05]
    dump.write_nuclide_case(tmp_path / "d.case", nuc, e_bins, params.scatt_type, params.order, params.mu_bins, True,
                            Ein, Ein_inel, params)
    for fmt, flag in (("binary", []), ("ascii", ["--ascii"])):
        cpp, py = tmp_path / f"cpp.{fmt}", tmp_path / f"py.{fmt}"
        r = subprocess.run([tool, str(tmp_path / "d.case"), str(tmp_path / "d.res"), "--library", str(cpp), "--name",
                            nuc.name, "--print-tol", "1e-8", "--thin-tol", "0.002"] + flag,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
        res = driver.preprocess_nuclide(nuc, e_bins, params, print_tol=1e-8, thin_tol=0.002, Ein_el=Ein,
                                        Ein_inel=Ein_inel, library_file=str(py), lib_format=fmt)
        assert len(res.Ein_el) <= len(Ein)
        assert open(cpp, "rb").read() == open(py, "rb").read()


@pytest.mark.gpu
def test_cpp_driver_builds_the_ein_grids_on_the_device(tool, tmp_path):
    """--ein-grid: the per-nuclide body of preprocess_ndpp including create_Ein_grid (src/scatt.F90:139) from the C++ host
    layer (ScattDataSet::create_Ein_grid) against the Python driver with no grids given: identical library files, and the
    grid in the file is the oracle's."""
    from ndpp_b200 import driver, output
    from oracle import pyoracle
    nuc = small_heavy()
    nuc.name = "92238.71c"
    e_bins = synth.group_structure(70)
    params = ace.Params(order=3, mu_bins=501, nuscatter=False)
    dump.write_nuclide_case(tmp_path / "g.case", nuc, e_bins, params.scatt_type, params.order, params.mu_bins, False,
                            np.array([1.0]), np.array([1.0]), params)
    cpp, py = tmp_path / "cpp.bin", tmp_path / "py.bin"
    r = subprocess.run([tool, str(tmp_path / "g.case"), str(tmp_path / "g.res"), "--library", str(cpp), "--name", nuc.name,
                        "--print-tol", "1e-8", "--thin-tol", "0", "--ein-grid"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    res = driver.preprocess_nuclide(nuc, e_bins, params, print_tol=1e-8, thin_tol=0.0, library_file=str(py))
    assert open(cpp, "rb").read() == open(py, "rb").read()
    rn = pyoracle.RefNuclide(nuc, e_bins, params)
    el, inel = rn.create_ein_grid()
    rn.close()
    assert np.array_equal(res.Ein_el, el) and np.array_equal(res.Ein_inel, inel)


@pytest.mark.gpu
def test_cpp_calc_scatt_c1(tool, oracle, tmp_path):
    from ndpp_b200 import scatt
    nuc, e_bins, params = synth.c1_fixture()
    Ein = synth.c1_ein_grid(40)
    Ein_inel = Ein[Ein >= 2.0]
    dump.write_nuclide_case(tmp_path / "c1.case", nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, True, Ein, Ein_inel,
                            params)
    r = _run(tool, tmp_path / "c1.case", tmp_path / "c1.res")
    assert "moment evaluations" in r.stdout
    el, inel, nu = dump.read_result(tmp_path / "c1.res")
    pel, pinel, pnu = scatt.calc_scatt(nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, True, Ein, Ein_inel)
    assert np.array_equal(el, pel) and np.array_equal(inel, pinel) and np.array_equal(nu, pnu)
    rn = oracle.RefNuclide(nuc, e_bins, ace.Params(order=5, mu_bins=3001, nuscatter=True))
    rn.convert_distro()
    assert_parity(el, rn.elastic(Ein), what="C++ calc_scatt C1 elastic")
    ri, rnu = rn.inelastic(Ein_inel)
    assert_parity(inel, ri, what="C++ calc_scatt C1 inelastic")
    assert_parity(nu, rnu, what="C++ calc_scatt C1 nu-inelastic")
    # nuscatt = .false. and no inelastic grid: the two matrices stay unallocated (src/scatt.F90:146-150)
    dump.write_nuclide_case(tmp_path / "c1b.case", nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, False, Ein, None, params)
    _run(tool, tmp_path / "c1b.case", tmp_path / "c1b.res")
    el2, inel2, nu2 = dump.read_result(tmp_path / "c1b.res")
    assert inel2 is None and nu2 is None and np.array_equal(el2, el)


@pytest.mark.gpu
def test_cpp_calc_scatt_heavy_shape(tool, oracle, tmp_path):
    nuc = small_heavy()
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7, mu_bins=2001, nuscatter=True)
    rng = np.random.default_rng(11)
    Ein = np.sort(np.exp(rng.uniform(np.log(1e-10), np.log(19.9), 60)))
    Ein_inel = Ein[Ein >= 0.05]
    dump.write_nuclide_case(tmp_path / "h.case", nuc, e_bins, params.scatt_type, params.order, params.mu_bins, True,
                            Ein, Ein_inel, params)
    _run(tool, tmp_path / "h.case", tmp_path / "h.res")
    el, inel, nu = dump.read_result(tmp_path / "h.res")
    rn = oracle.RefNuclide(nuc, e_bins, params)
    rn.convert_distro()
    assert_parity(el, rn.elastic(Ein), what="C++ calc_scatt heavy elastic")
    ri, rnu = rn.inelastic(Ein_inel)
    assert_parity(inel, ri, what="C++ calc_scatt heavy inelastic")
    assert_parity(nu, rnu, what="C++ calc_scatt heavy nu-inelastic")


@pytest.mark.gpu
@pytest.mark.parametrize("mode,elastic", [("skewed", None), ("cont", "incoherent")])
def test_cpp_calc_scattsab(tool, oracle, tmp_path, mode, elastic):
    sab = synth.c4_sab(mode=mode, elastic=elastic, n_ein=30)
    e_bins = synth.group_structure(70)
    rng = np.random.default_rng(5)
    Ein = np.sort(np.concatenate([sab.inelastic_e_in, np.exp(rng.uniform(np.log(1e-11), np.log(4e-6), 100)), [5e-6]]))
    dump.write_sab_case(tmp_path / "s.case", sab, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 2001, Ein)
    _run(tool, tmp_path / "s.case", tmp_path / "s.res")
    got, _, _ = dump.read_result(tmp_path / "s.res")
    assert_parity(got, oracle.sab_calc(sab, e_bins, 5, Ein), what=f"C++ calc_scattsab {mode}")


@pytest.mark.gpu
def test_cpp_fatal_error_carries_the_reference_message(tool, tmp_path):
    nuc, e_bins, params = synth.c1_fixture()
    # E_in above the last tabulated adist energy: the reference aborts in binary_search (src/search.F90:36-38)
    dump.write_nuclide_case(tmp_path / "e.case", nuc, e_bins, ace.SCATT_TYPE_LEGENDRE, 5, 3001, False,
                            np.array([1.5]), np.array([2.7]), params)
    r = _run(tool, tmp_path / "e.case", tmp_path / "e.res", expect_ok=False)
    assert r.returncode != 0 and r.stderr.startswith(" ERROR: ") and "binary search" in r.stderr
    assert not os.path.exists(tmp_path / "e.res")


@pytest.mark.gpu
def test_cpp_calc_scatt_on_a_device_group(tool, tmp_path):
    """`ndpp_calc_scatt --devices N` = calc_scatt on every GPU of the box through ndpp_host::DeviceGroup (ndppgpu_group_*:
    E_in dealt cyclically over the devices, NCCL gather to device 0); its matrices equal the one-device program's bit for
    bit, with and without the library step (tolerance + thinning on the root device)."""
    import torch
    nuc = small_heavy()
    e_bins = synth.group_structure(70)
    params = ace.Params(order=7, mu_bins=2001, nuscatter=True)
    rng = np.random.default_rng(12)
    Ein = np.sort(np.concatenate([np.exp(rng.uniform(np.log(1e-10), np.log(19.9), 90)), [20.0, 20.01]]))
    Ein_inel = Ein[Ein >= 0.05]
    dump.write_nuclide_case(tmp_path / "g.case", nuc, e_bins, params.scatt_type, params.order, params.mu_bins, True,
                            Ein, Ein_inel, params)
    _run(tool, tmp_path / "g.case", tmp_path / "one.res")
    n = torch.cuda.device_count()
    r = subprocess.run([tool, str(tmp_path / "g.case"), str(tmp_path / "grp.res"), "--devices", "0"], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, r.stderr
    assert f" {n} devices:" in r.stdout
    a, b = dump.read_result(tmp_path / "one.res"), dump.read_result(tmp_path / "grp.res")
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    for extra, tag in ((["--device", "0"], "one"), (["--devices", "0"], "grp")):
        r = subprocess.run([tool, str(tmp_path / "g.case"), str(tmp_path / f"{tag}2.res"), "--library",
                            str(tmp_path / f"{tag}.lib"), "--print-tol", "1e-8", "--thin-tol", "0.002"] + extra,
                           capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr
    assert open(tmp_path / "one.lib", "rb").read() == open(tmp_path / "grp.lib", "rb").read()
    r = subprocess.run([tool, str(tmp_path / "g.case"), str(tmp_path / "x.res"), "--devices", str(n + 1)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode != 0 and "ERROR: ndppgpu_group_init: more devices requested than the box has" in r.stderr
