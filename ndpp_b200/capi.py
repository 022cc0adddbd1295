"""ctypes binding of the C-ABI in include/ndppgpu.h (libndppgpu.so).

This is the Python twin of the ISO_C_BINDING interface module the reference's Fortran driver would
use.  Loading fails loudly when the library has not been built, and every entry
point raises NdppGpuError on a non-zero status -- there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libndppgpu.so")

c_dp = C.POINTER(C.c_double)
c_ip = C.POINTER(C.c_int)

# every symbol include/ndppgpu.h declares
EXPORTS = ["ndppgpu_init", "ndppgpu_finalize", "ndppgpu_last_error", "ndppgpu_stats", "ndppgpu_abi_version",
           "ndppgpu_stream", "ndppgpu_nuclide_create", "ndppgpu_nuclide_add_reaction", "ndppgpu_convert_distro",
           "ndppgpu_elastic", "ndppgpu_inelastic", "ndppgpu_elastic_dev", "ndppgpu_inelastic_dev",
           "ndppgpu_nuclide_n_slots", "ndppgpu_nuclide_slot_info", "ndppgpu_nuclide_slot_row_np",
           "ndppgpu_nuclide_get_table", "ndppgpu_nuclide_free", "ndppgpu_sab_create", "ndppgpu_sab", "ndppgpu_sab_dev",
           "ndppgpu_sab_free", "ndppgpu_measure_fp64_peak", "ndppgpu_interp_distro", "ndppgpu_test_legendre", "ndppgpu_nuclide_set_table",
           "ndppgpu_test_exact_math", "ndppgpu_apply_tol", "ndppgpu_apply_tol_dev", "ndppgpu_thin_grid",
           "ndppgpu_thin_grid_dev", "ndppgpu_gather_columns_dev", "ndppgpu_elastic_thinned",
           "ndppgpu_inelastic_thinned", "ndppgpu_chi", "ndppgpu_eval_libm",
           "ndppgpu_calc_scatt", "ndppgpu_nuclide_create_ein_grid", "ndppgpu_nuclide_ein_grid", "ndppgpu_sab_egrid", "ndppgpu_sab_ein_grid",
           "ndppgpu_group_init", "ndppgpu_group_unique_id", "ndppgpu_group_init_rank", "ndppgpu_group_info",
           "ndppgpu_group_ctx", "ndppgpu_group_gathered_bytes", "ndppgpu_group_finalize",
           "ndppgpu_group_nuclide_create", "ndppgpu_group_nuclide_add_reaction", "ndppgpu_group_convert_distro",
           "ndppgpu_group_elastic", "ndppgpu_group_inelastic", "ndppgpu_group_nuclide_free", "ndppgpu_group_set_grids",
           "ndppgpu_group_integrate", "ndppgpu_group_sync", "ndppgpu_group_join", "ndppgpu_group_fetch", "ndppgpu_group_result_dev",
           "ndppgpu_plan_library", "ndppgpu_tile_bounds", "ndppgpu_library_create", "ndppgpu_library_run",
           "ndppgpu_library_fetch", "ndppgpu_library_report_get", "ndppgpu_library_free"]


class NdppGpuError(RuntimeError):
    pass


class ParamsC(C.Structure):
    _fields_ = [("scatt_type", C.c_int), ("order", C.c_int), ("mu_bins", C.c_int), ("nuscatter", C.c_int),
                ("ne_per_grp", C.c_int), ("adaptive_mu_its", C.c_int), ("adaptive_eout_its", C.c_int),
                ("reserved", C.c_int), ("sab_threshold", C.c_double), ("brent_mu_thresh", C.c_double),
                ("adaptive_mu_tol", C.c_double), ("adaptive_eout_tol", C.c_double)]


class StatsC(C.Structure):
    _fields_ = [("kernel_ms", C.c_double), ("h2d_bytes", C.c_double), ("d2h_bytes", C.c_double),
                ("launches", C.c_longlong), ("moment_evals", C.c_longlong), ("file4_calls", C.c_longlong),
                ("file6_cm_points", C.c_longlong), ("file6_lab_calls", C.c_longlong), ("freegas_tasks", C.c_longlong),
                ("sab_columns", C.c_longlong), ("file6_cm_ms", C.c_double), ("file6_cm_launches", C.c_longlong),
                ("host_call_ms", C.c_double), ("host_alloc_ms", C.c_double), ("host_sync_ms", C.c_double),
                ("freegas_items", C.c_longlong), ("freegas_kernel_evals", C.c_longlong), ("freegas_sab_evals", C.c_longlong)]


class ShapeC(C.Structure):
    _fields_ = [("index", C.c_int), ("n_el", C.c_int), ("n_inel", C.c_int), ("n_levels", C.c_int), ("has_cont", C.c_int),
                ("freegas_points", C.c_int), ("cont_threshold", C.c_double), ("e_lo", C.c_double), ("e_hi", C.c_double),
                ("level_thresholds", C.POINTER(C.c_double))]


class ItemC(C.Structure):
    _fields_ = [("nuclide", C.c_int), ("matrix", C.c_int), ("tile", C.c_int), ("n_tiles", C.c_int), ("rank", C.c_int),
                ("rows", C.c_int), ("cost", C.c_double)]


class LibraryReportC(C.Structure):
    _fields_ = [("wall_s", C.c_double), ("compute_s", C.c_double), ("gather_s", C.c_double), ("open_s_max", C.c_double),
                ("integrate_s_max", C.c_double), ("kernel_s_max", C.c_double), ("kernel_s_sum", C.c_double),
                ("moment_evals", C.c_longlong), ("items", C.c_int), ("opens", C.c_int), ("device_s_max", C.c_double),
                ("alloc_s_max", C.c_double), ("reserved", C.c_double * 2)]


OPEN_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_double)),
                      C.POINTER(C.c_int), C.POINTER(C.POINTER(C.c_double)), C.POINTER(C.c_int))
CLOSE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.c_void_p)

_lib = None


def load() -> C.CDLL:
    """Load libndppgpu.so; raise if it is missing (build with `python -m ndpp_b200.build`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise NdppGpuError(f"{SO_PATH} not found: build the CUDA library first (python ndpp_b200/build.py). "
                           "ndpp_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    L.ndppgpu_init.argtypes = [i, C.POINTER(vp)]
    L.ndppgpu_finalize.argtypes = [vp]
    L.ndppgpu_last_error.argtypes = [vp, C.c_char_p, i]
    L.ndppgpu_stats.argtypes = [vp, C.POINTER(StatsC), i]
    L.ndppgpu_stream.argtypes = [vp]
    L.ndppgpu_stream.restype = vp
    L.ndppgpu_nuclide_create.argtypes = [vp, d, d, d, i, c_dp, c_dp, c_dp, i, C.POINTER(ParamsC), C.POINTER(vp)]
    L.ndppgpu_nuclide_add_reaction.argtypes = [vp, i, i, d, i, i, i, i, i, i, c_dp, i, c_dp, i, c_dp, i, c_dp, c_ip,
                                               c_ip, i, c_dp, i, c_dp, i]
    L.ndppgpu_convert_distro.argtypes = [vp]
    L.ndppgpu_elastic.argtypes = [vp, c_dp, i, c_dp]
    L.ndppgpu_inelastic.argtypes = [vp, c_dp, i, c_dp, c_dp]
    L.ndppgpu_calc_scatt.argtypes = [vp, c_dp, i, c_dp, c_dp, i, c_dp, c_dp]
    L.ndppgpu_elastic_dev.argtypes = [vp, vp, i, vp]
    L.ndppgpu_inelastic_dev.argtypes = [vp, vp, i, vp, vp]
    L.ndppgpu_nuclide_n_slots.argtypes = [vp]
    L.ndppgpu_nuclide_slot_info.argtypes = [vp, i, c_ip]
    L.ndppgpu_nuclide_slot_row_np.argtypes = [vp, i, i]
    L.ndppgpu_nuclide_get_table.argtypes = [vp, i, i, c_dp, c_dp, c_dp, c_dp, c_ip]
    L.ndppgpu_nuclide_free.argtypes = [vp]
    L.ndppgpu_nuclide_create_ein_grid.argtypes = [vp, i, i, c_ip, c_ip, c_ip]
    L.ndppgpu_nuclide_ein_grid.argtypes = [vp, i, c_dp, C.POINTER(vp)]
    L.ndppgpu_sab_egrid.argtypes = [vp, c_dp, i, i, i, c_ip, c_ip]
    L.ndppgpu_sab_ein_grid.argtypes = [vp, c_dp, C.POINTER(vp)]
    L.ndppgpu_sab_create.argtypes = [vp, d, d, d, d, i, i, i, i, c_dp, c_dp, c_dp, c_dp, c_ip, c_dp, c_dp, c_dp, i,
                                     i, i, c_dp, c_dp, c_dp, C.POINTER(vp)]
    L.ndppgpu_sab.argtypes = [vp, c_dp, i, i, i, c_dp, i, c_dp, c_dp, c_dp]
    L.ndppgpu_sab_dev.argtypes = [vp, c_dp, i, i, i, vp, i, vp]
    L.ndppgpu_sab_free.argtypes = [vp]
    L.ndppgpu_measure_fp64_peak.argtypes = [vp, d, C.POINTER(d)]
    L.ndppgpu_chi.argtypes = [vp, i, c_dp, c_dp, i, c_dp, i, i, c_dp, i, i, c_dp, i, i, vp, c_dp, i, c_dp, i, c_dp, i,
                              c_dp, c_dp, c_dp]
    L.ndppgpu_interp_distro.argtypes = [vp, i, c_dp, i, c_dp]
    L.ndppgpu_nuclide_set_table.argtypes = [vp, i, i, c_dp]
    L.ndppgpu_test_legendre.argtypes = [vp, i, i, c_dp, c_dp, c_dp, c_dp, c_dp, c_dp]
    L.ndppgpu_eval_libm.argtypes = [vp, i, c_dp, C.c_longlong, c_dp]
    L.ndppgpu_test_exact_math.argtypes = [vp, C.c_ulonglong, i, C.POINTER(C.c_ulonglong)]
    L.ndppgpu_apply_tol.argtypes = [vp, c_dp, i, i, i, d]
    L.ndppgpu_apply_tol_dev.argtypes = [vp, vp, i, i, i, d]
    L.ndppgpu_thin_grid.argtypes = [vp, c_dp, c_dp, c_dp, i, i, c_dp, i, d, c_ip, c_dp, c_dp]
    L.ndppgpu_thin_grid_dev.argtypes = [vp, vp, vp, vp, i, i, c_dp, i, d, vp, c_ip, c_dp, c_dp]
    L.ndppgpu_gather_columns_dev.argtypes = [vp, vp, vp, i, i, vp]
    L.ndppgpu_elastic_thinned.argtypes = [vp, c_dp, i, d, d, c_dp, i, c_dp, c_ip, c_dp, c_dp]
    L.ndppgpu_inelastic_thinned.argtypes = [vp, c_dp, i, d, d, c_dp, i, c_dp, c_dp, c_ip, c_dp, c_dp]
    # several GPUs (csrc/group.cuh)
    L.ndppgpu_group_init.argtypes = [i, c_ip, C.POINTER(vp)]
    L.ndppgpu_group_unique_id.argtypes = [vp]
    L.ndppgpu_group_init_rank.argtypes = [i, i, i, vp, C.POINTER(vp)]
    L.ndppgpu_group_info.argtypes = [vp, c_ip, c_ip, c_ip]
    L.ndppgpu_group_ctx.argtypes = [vp, i]
    L.ndppgpu_group_ctx.restype = vp
    L.ndppgpu_group_gathered_bytes.argtypes = [vp, i]
    L.ndppgpu_group_gathered_bytes.restype = C.c_longlong
    L.ndppgpu_group_finalize.argtypes = [vp]
    L.ndppgpu_group_nuclide_create.argtypes = [vp, d, d, d, i, c_dp, c_dp, c_dp, i, C.POINTER(ParamsC), C.POINTER(vp)]
    L.ndppgpu_group_nuclide_add_reaction.argtypes = L.ndppgpu_nuclide_add_reaction.argtypes
    L.ndppgpu_group_convert_distro.argtypes = [vp]
    L.ndppgpu_group_elastic.argtypes = [vp, c_dp, i, c_dp]
    L.ndppgpu_group_inelastic.argtypes = [vp, c_dp, i, c_dp, c_dp]
    L.ndppgpu_group_nuclide_free.argtypes = [vp]
    L.ndppgpu_group_set_grids.argtypes = [vp, c_dp, i, c_dp, i]
    L.ndppgpu_group_integrate.argtypes = [vp, i]
    L.ndppgpu_group_sync.argtypes = [vp]
    L.ndppgpu_group_join.argtypes = [vp]
    L.ndppgpu_group_fetch.argtypes = [vp, c_dp, c_dp, c_dp]
    L.ndppgpu_group_result_dev.argtypes = [vp, i]
    L.ndppgpu_group_result_dev.restype = vp
    L.ndppgpu_plan_library.argtypes = [C.POINTER(ShapeC), i, i, i, i, i, i, i, d, i, C.POINTER(ItemC), i, c_ip, c_dp]
    L.ndppgpu_tile_bounds.argtypes = [i, i, i, c_ip, c_ip]
    L.ndppgpu_tile_bounds.restype = None
    L.ndppgpu_library_create.argtypes = [vp, i, i, i, C.POINTER(ItemC), i, C.POINTER(vp)]
    L.ndppgpu_library_run.argtypes = [vp, OPEN_FN, CLOSE_FN, vp, C.POINTER(LibraryReportC)]
    L.ndppgpu_library_fetch.argtypes = [vp, i, i, c_dp, i, d, c_dp]
    L.ndppgpu_library_report_get.argtypes = [vp, C.POINTER(LibraryReportC)]
    L.ndppgpu_library_free.argtypes = [vp]
    _lib = L
    return L


def f64(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def i32(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())


def dp(a):
    return None if a is None else a.ctypes.data_as(c_dp)


def ip(a):
    return None if a is None else a.ctypes.data_as(c_ip)


def make_params(p) -> ParamsC:
    return ParamsC(int(p.scatt_type), int(p.order), int(p.mu_bins), int(bool(p.nuscatter)), int(p.ne_per_grp),
                   int(p.adaptive_mu_its), int(p.adaptive_eout_its), 0, float(p.sab_threshold),
                   float(p.brent_mu_thresh), float(p.adaptive_mu_tol), float(p.adaptive_eout_tol))


def last_error(ctx=None) -> str:
    buf = C.create_string_buffer(1024)
    load().ndppgpu_last_error(ctx, buf, 1024)
    return buf.value.decode(errors="replace")


def check(rc, ctx=None):
    if rc != 0:
        raise NdppGpuError(last_error(ctx))


class Context:
    """One device context (ndppgpu_init / ndppgpu_finalize)."""

    def __init__(self, device: int = -1, borrowed=None):
        self.lib = load()
        self.h = C.c_void_p()
        self.owned = borrowed is None
        if borrowed is not None:           # a context that belongs to a device group (ndppgpu_group_ctx)
            self.h = C.c_void_p(borrowed)
        else:
            check(self.lib.ndppgpu_init(int(device), C.byref(self.h)))

    def stats(self, reset=False) -> dict:
        s = StatsC()
        check(self.lib.ndppgpu_stats(self.h, C.byref(s), int(reset)), self.h)
        return {k: getattr(s, k) for k, _ in StatsC._fields_}

    def measure_fp64_peak(self, seconds=0.5) -> float:
        out = C.c_double(0.0)
        check(self.lib.ndppgpu_measure_fp64_peak(self.h, float(seconds), C.byref(out)), self.h)
        return out.value

    def test_legendre(self, L, xlow, xhigh, flow, fhigh):
        xl, xh, fl, fh = f64(xlow), f64(xhigh), f64(flow), f64(fhigh)
        integ, pn = np.empty((len(xl), L)), np.empty((len(xl), L))
        check(self.lib.ndppgpu_test_legendre(self.h, len(xl), L, dp(xl), dp(xh), dp(fl), dp(fh), dp(integ), dp(pn)),
              self.h)
        return integ, pn

    def test_exact_math(self, seed=1, per_thread=2000):
        out = (C.c_ulonglong * 2)()
        check(self.lib.ndppgpu_test_exact_math(self.h, int(seed), int(per_thread), out), self.h)
        return {"pairs": out[0], "mismatch": out[1]}

    def eval_libm(self, fn, x):
        """The device's exp (0) / expm1 (1) / sinh (2) / cosh (3) of csrc/libm_exact.cuh at the host array x."""
        x = f64(x)
        y = np.empty_like(x)
        check(self.lib.ndppgpu_eval_libm(self.h, int(fn), dp(x), len(x), dp(y)), self.h)
        return y

    @property
    def stream(self) -> int:
        return int(self.lib.ndppgpu_stream(self.h) or 0)

    def close(self):
        if self.h:
            if self.owned:
                self.lib.ndppgpu_finalize(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
