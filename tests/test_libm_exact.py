"""sinh / cosh with the host libm's bits (ndpp_b200/csrc/libm_exact.cuh).

The reference's Law 44 conversion (src/scattdata_header.F90:822-831) calls the C library's sinh / cosh; the closed
forms downstream amplify a last-bit change of a table value to ~1e-12 on a moment, so the CUDA path restates the
algorithms GNU libc 2.39 runs (x86-64, FMA builds of exp / expm1).  Bit equality with the *running* libm is measured
here: on the host build of the header (CPU test, >= 1e8 arguments per function) and on the device (GPU test, >= 1e8
arguments of sinh and cosh over the Law-44 range |A mu| <= 40 and beyond).
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "native", "libm_exact_check.cpp")
SO = os.path.join(HERE, "native", "libm_exact_check.so")
FN = {"exp": 0, "expm1": 1, "sinh": 2, "cosh": 3, "log": 4}


@pytest.fixture(scope="module")
def chk():
    deps = [SRC, os.path.join(HERE, "..", "ndpp_b200", "csrc", "libm_exact.cuh"),
            os.path.join(HERE, "..", "ndpp_b200", "csrc", "exp_table.inc"),
            os.path.join(HERE, "..", "ndpp_b200", "csrc", "log_table.inc")]
    if not os.path.exists(SO) or any(os.path.getmtime(d) > os.path.getmtime(SO) for d in deps):
        subprocess.check_call(["g++", "-O2", "-mfma", "-ffp-contract=off", "-fopenmp", "-shared", "-fPIC", "-o", SO, SRC])
    L = C.CDLL(SO)
    L.libm_exact_mismatches.restype = C.c_longlong
    L.libm_exact_mismatches.argtypes = [C.c_int, C.c_double, C.c_double, C.c_longlong, C.c_ulonglong, C.c_int,
                                        C.POINTER(C.c_double)]
    dp = C.POINTER(C.c_double)
    L.libm_host_eval.argtypes = [C.c_int, dp, dp, C.c_longlong]
    L.libm_port_eval.argtypes = [C.c_int, dp, dp, C.c_longlong]
    return L


def _has_fma():
    try:
        return " fma " in open("/proc/cpuinfo").read()
    except OSError:
        return False


# (lo, hi, mode): uniform in [lo, hi], or sign * 10^uniform(lo, hi)
RANGES = [(-40.0, 40.0, 0), (-1.0, 1.0, 0), (-0.7, 0.7, 0), (-710.0, 710.0, 0), (-320.0, 3.0, 1)]


@pytest.mark.skipif(not _has_fma(), reason="the restated libm copies are the ones glibc selects on FMA + AVX2 CPUs")
@pytest.mark.parametrize("name", ["exp", "expm1", "sinh", "cosh"])
def test_host_build_reproduces_the_running_libm(chk, name):
    total = 0
    for k, (lo, hi, mode) in enumerate(RANGES):
        n = 24_000_000
        bad_x = C.c_double(0.0)
        bad = chk.libm_exact_mismatches(FN[name], lo, hi, n, 20261018 + 16 * FN[name] + k, mode, C.byref(bad_x))
        assert bad == 0, f"{name} on [{lo}, {hi}] mode {mode}: {bad} of {n} results differ from libm, e.g. at x = {bad_x.value!r}"
        total += n
    assert total >= 100_000_000


# log: the ratios the grid builders pass (Ehi / Elo of neighbouring points and of group edges: 1 .. 1e12), the window
# around 1 with its own polynomial, every binade, subnormal arguments
LOG_RANGES = [(0.9, 1.1, 0), (0.93, 1.07, 0), (0.0, 4.0, 0), (1.0, 1.0e6, 0), (0.0, 12.0, 1), (-320.0, 308.0, 1)]


@pytest.mark.skipif(not _has_fma(), reason="the restated libm copies are the ones glibc selects on FMA + AVX2 CPUs")
def test_host_build_of_log_reproduces_the_running_libm(chk):
    total = 0
    for k, (lo, hi, mode) in enumerate(LOG_RANGES):
        n = 24_000_000
        bad_x = C.c_double(0.0)
        bad = chk.libm_exact_mismatches(FN["log"], lo, hi, n, 20261019 + k, mode, C.byref(bad_x))
        assert bad == 0, f"log on [{lo}, {hi}] mode {mode}: {bad} of {n} results differ from libm, e.g. at x = {bad_x.value!r}"
        total += n
    assert total >= 100_000_000


def test_special_arguments_of_log(chk):
    x = np.array([0.0, -0.0, 5e-324, 1e-320, 2.2250738585072014e-308, 1.0, np.nextafter(1.0, 0), np.nextafter(1.0, 2),
                  1.0 - 2.0 ** -4, np.nextafter(1.0 - 2.0 ** -4, 0), 1.0 + 265.0 / 4096.0, np.nextafter(1.0 + 265.0 / 4096.0, 0), 2.0,
                  1e308, 1.7976931348623157e308, np.inf, -1.0, -np.inf, np.nan])
    a, b = np.empty_like(x), np.empty_like(x)
    dp = C.POINTER(C.c_double)
    with np.errstate(all="ignore"):
        chk.libm_port_eval(FN["log"], x.ctypes.data_as(dp), a.ctypes.data_as(dp), len(x))
        chk.libm_host_eval(FN["log"], x.ctypes.data_as(dp), b.ctypes.data_as(dp), len(x))
    same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
    assert same.all(), (x[~same], a[~same], b[~same])


@pytest.mark.gpu
def test_device_log_carries_the_host_libm_bits(chk):
    from ndpp_b200 import scatt
    ctx = scatt.default_context()
    rng = np.random.default_rng(404)
    dp = C.POINTER(C.c_double)
    total = 0
    for lo, hi, mode, reps in ((0.9, 1.1, 0, 2), (0.0, 4.0, 0, 2), (1.0, 1.0e6, 0, 2), (0.0, 12.0, 1, 2), (-320.0, 308.0, 1, 2)):
        for _ in range(reps):
            n = 10_000_000
            x = rng.uniform(lo, hi, n)
            if mode == 1:
                x = 10.0 ** x
            ref = np.empty_like(x)
            chk.libm_host_eval(FN["log"], x.ctypes.data_as(dp), ref.ctypes.data_as(dp), n)
            got = ctx.eval_libm(FN["log"], x)
            bad = np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0]
            assert bad.size == 0, f"log: {bad.size} of {n} device results differ from libm, e.g. x = {x[bad[0]]!r}"
            total += n
    assert total >= 100_000_000


def test_special_arguments(chk):
    x = np.array([0.0, -0.0, 1e-320, -1e-320, 2.0 ** -54, 2.0 ** -28, 0.5 * np.log(2.0), 1.5 * np.log(2.0), 1.0, 22.0,
                  -22.0, 56 * np.log(2.0), 709.0, 709.78, 710.0, 710.4758600739439, 710.5, 745.0, -745.2, 1e300, np.inf,
                  -np.inf, np.nan, np.nextafter(22.0, 0), np.nextafter(1.0, 0), 0.34657359027997264, 1.0397207708399179])
    x = np.concatenate([x, -x])
    a, b = np.empty_like(x), np.empty_like(x)
    dp = C.POINTER(C.c_double)
    for name in ("exp", "expm1", "sinh", "cosh"):
        with np.errstate(all="ignore"):
            chk.libm_port_eval(FN[name], x.ctypes.data_as(dp), a.ctypes.data_as(dp), len(x))
            chk.libm_host_eval(FN[name], x.ctypes.data_as(dp), b.ctypes.data_as(dp), len(x))
        same = (a.view(np.uint64) == b.view(np.uint64)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), (name, x[~same], a[~same], b[~same])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["sinh", "cosh"])
def test_device_sinh_cosh_carry_the_host_libm_bits(chk, name):
    """>= 1e8 arguments per function: the Law-44 range (A up to ~40, mu in [-1, 1]), the small-argument branches, and the
    exp-only branch up to the overflow threshold."""
    from ndpp_b200 import scatt
    ctx = scatt.default_context()
    rng = np.random.default_rng(44 + FN[name])
    dp = C.POINTER(C.c_double)
    total = 0
    for lo, hi, mode, reps in ((-40.0, 40.0, 0, 5), (-1.0, 1.0, 0, 3), (-710.4, 710.4, 0, 1), (-30.0, 2.0, 1, 2)):
        for _ in range(reps):
            n = 10_000_000
            x = rng.uniform(lo, hi, n)
            if mode == 1:
                x = np.where(rng.random(n) < 0.5, -1.0, 1.0) * 10.0 ** x
            ref = np.empty_like(x)
            chk.libm_host_eval(FN[name], x.ctypes.data_as(dp), ref.ctypes.data_as(dp), n)
            got = ctx.eval_libm(FN[name], x)
            bad = np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0]
            assert bad.size == 0, f"{name}: {bad.size} of {n} device results differ from libm, e.g. x = {x[bad[0]]!r}: " \
                                  f"{got[bad[0]]!r} vs {ref[bad[0]]!r}"
            total += n
    assert total >= 100_000_000


@pytest.mark.gpu
def test_device_exp_expm1(chk):
    from ndpp_b200 import scatt
    ctx = scatt.default_context()
    rng = np.random.default_rng(7)
    dp = C.POINTER(C.c_double)
    for name in ("exp", "expm1"):
        x = np.concatenate([rng.uniform(-745.0, 710.0, 4_000_000), rng.uniform(-2.0, 2.0, 4_000_000)])
        ref = np.empty_like(x)
        chk.libm_host_eval(FN[name], x.ctypes.data_as(dp), ref.ctypes.data_as(dp), len(x))
        got = ctx.eval_libm(FN[name], x)
        assert np.array_equal(got.view(np.uint64), ref.view(np.uint64)), name


@pytest.mark.gpu
def test_device_free_gas_primitives_carry_the_library_bits(chk):
    """The branch-free division, square root and exp of the free-gas kernel's two-point evaluation (fg_div_fast,
    fg_sqrt_fast, fg_exp_neg in csrc/kernels_freegas.cuh) against `/`, sqrt and the host libm's exp: >= 1e8 arguments
    each over the magnitudes the kernel meets (and far beyond), bit for bit wherever their range guard holds (elsewhere
    the device returns the library operation itself, ndppgpu.h: ndppgpu_eval_libm fn 10-12)."""
    from ndpp_b200 import scatt
    ctx = scatt.default_context()
    rng = np.random.default_rng(1012)
    dp = C.POINTER(C.c_double)
    n = 10_000_000
    totals = [0, 0, 0]
    for rep in range(10):
        # exp on (-708, 0]: uniform, and log-uniform magnitudes down to the tiny branch
        x = -rng.uniform(0.0, 708.0, n) if rep < 6 else -(10.0 ** rng.uniform(-20.0, 2.85, n))
        x = np.maximum(x, -707.999)
        ref = np.empty_like(x)
        chk.libm_host_eval(FN["exp"], x.ctypes.data_as(dp), ref.ctypes.data_as(dp), n)
        got = ctx.eval_libm(10, x)
        bad = np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, f"exp: {bad.size} of {n}, e.g. x = {x[bad[0]]!r}: {got[bad[0]]!r} vs {ref[bad[0]]!r}"
        totals[0] += n
        # sqrt: mantissas uniform, exponents over the whole range (and the kernel's 4 pi alpha in [1e-5, 1e6])
        x = rng.uniform(1.0, 4.0, n) * (2.0 ** rng.integers(-1020, 1020, n) if rep % 2 else 10.0 ** rng.uniform(-6, 7, n))
        got = ctx.eval_libm(11, x)
        ref = np.sqrt(x)
        bad = np.nonzero(got.view(np.uint64) != ref.view(np.uint64))[0]
        assert bad.size == 0, f"sqrt: {bad.size} of {n}, e.g. x = {x[bad[0]]!r}: {got[bad[0]]!r} vs {ref[bad[0]]!r}"
        totals[1] += n
        # division x[i] / x[i ^ 1]
        e = rng.uniform(-30, 30, n) if rep % 2 else rng.uniform(-290, 290, n)
        x = rng.uniform(1.0, 10.0, n) * 10.0 ** e * np.where(rng.random(n) < 0.3, -1.0, 1.0)
        x[::1000] = 0.0
        with np.errstate(all="ignore"):
            ref = x / x.reshape(-1, 2)[:, ::-1].reshape(-1)
        got = ctx.eval_libm(12, x)
        same = (got.view(np.uint64) == ref.view(np.uint64)) | (np.isnan(got) & np.isnan(ref))
        bad = np.nonzero(~same)[0]
        assert bad.size == 0, f"div: {bad.size} of {n}, e.g. {x[bad[0]]!r} / {x[bad[0] ^ 1]!r}: {got[bad[0]]!r} vs {ref[bad[0]]!r}"
        totals[2] += n
    assert min(totals) >= 100_000_000
