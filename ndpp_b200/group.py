"""Several GPUs: host-side mirror of the ndppgpu_group_* / ndppgpu_library_* entry points (csrc/group.cuh).

What the reference does in its own driver -- partition_work (static blocks of nuclides per MPI rank,
src/ndpp.F90:934-950), the nuclide loop (:549) and the hand-back of the results (:839-864) -- happens
inside libndppgpu.so: this module only marshals arguments, exactly as the ISO_C_BINDING shim of
INTEGRATION.md section 4 does.  Two ways to form a group, same calls afterwards:

  Group(n)                          every GPU of this process (host threads + ncclCommInitAll in the library)
  Group.from_rank(dev, rank, world, id)   one GPU per process (MPI ranks / torchrun); `id` comes from
                                    Group.unique_id() on rank 0, broadcast by the caller

There is no CPU fallback: without the library or without GPUs every call raises NdppGpuError.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

from . import capi
from .ace import SCATT_TYPE_LEGENDRE, Nuclide, Params
from .capi import Context, ItemC, LibraryReportC, NdppGpuError, ShapeC, check, dp, f64
from .scatt import reaction_args


class Group:
    def __init__(self, n_devices: int = 0, devices: Optional[Sequence[int]] = None, _handle=None):
        self.lib = capi.load()
        self.h = C.c_void_p()
        if _handle is not None:
            self.h = _handle
        else:
            dv = capi.i32(devices) if devices is not None else None
            check(self.lib.ndppgpu_group_init(int(n_devices), capi.ip(dv), C.byref(self.h)))
        w, n, f = C.c_int(0), C.c_int(0), C.c_int(0)
        check(self.lib.ndppgpu_group_info(self.h, C.byref(w), C.byref(n), C.byref(f)))
        self.world, self.n_local, self.first = w.value, n.value, f.value

    @staticmethod
    def unique_id() -> bytes:
        buf = (C.c_ubyte * 128)()
        check(capi.load().ndppgpu_group_unique_id(buf))
        return bytes(buf)

    @classmethod
    def from_rank(cls, device: int, rank: int, world: int, nccl_id: Optional[bytes]):
        lib = capi.load()
        h = C.c_void_p()
        buf = (C.c_ubyte * 128).from_buffer_copy(nccl_id) if nccl_id is not None else None
        check(lib.ndppgpu_group_init_rank(int(device), int(rank), int(world), buf, C.byref(h)))
        return cls(_handle=h)

    @property
    def is_root(self) -> bool:
        return self.first == 0

    def ctx(self, local_index: int = 0) -> Context:
        p = self.lib.ndppgpu_group_ctx(self.h, int(local_index))
        if not p:
            raise NdppGpuError("ndppgpu_group_ctx: no such local device")
        return Context(borrowed=p)

    def stats(self, reset=False) -> List[dict]:
        return [self.ctx(i).stats(reset) for i in range(self.n_local)]

    def gathered_bytes(self, reset=False) -> int:
        return int(self.lib.ndppgpu_group_gathered_bytes(self.h, int(reset)))

    def err(self) -> str:
        return capi.last_error(self.lib.ndppgpu_group_ctx(self.h, 0))

    def close(self):
        if self.h:
            self.lib.ndppgpu_group_finalize(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class GroupNuclide:
    """DeviceNuclide on a group: the nuclide replicated on every device, E_in dealt cyclically, the columns gathered to
    the root over NCCL (csrc/group.cuh).  Results are identical to the one-device call bit for bit."""

    def __init__(self, nuc: Nuclide, energy_bins, params: Params, group: Group, convert=True):
        self.group, self.lib, self.params = group, group.lib, params
        self.e_bins = f64(energy_bins)
        self.G = len(self.e_bins) - 1
        self.L = params.order + 1 if params.scatt_type == SCATT_TYPE_LEGENDRE else params.order
        self.h = C.c_void_p()
        self._ctx0 = self.lib.ndppgpu_group_ctx(group.h, 0)
        en, el = f64(nuc.energy), f64(nuc.elastic)
        pc = capi.make_params(params)
        self._check(self.lib.ndppgpu_group_nuclide_create(group.h, nuc.awr, nuc.kT, nuc.freegas_cutoff, len(en), dp(en),
                                                          dp(el), dp(self.e_bins), len(self.e_bins), C.byref(pc),
                                                          C.byref(self.h)))
        for args in reaction_args(nuc):
            self._check(self.lib.ndppgpu_group_nuclide_add_reaction(self.h, *args))
        self.n_el = self.n_inel = 0
        if convert:
            self.convert_distro()

    def _check(self, rc):
        check(rc, self._ctx0)

    def convert_distro(self):
        self._check(self.lib.ndppgpu_group_convert_distro(self.h))

    # -- the seam replacements, host buffers (root receives) --------------------------------------------------------
    def _out(self, out, n):
        if not self.group.is_root:
            return None
        if out is None:
            return np.empty((n, self.G, self.L))
        assert out.dtype == np.float64 and out.flags.c_contiguous and out.size == n * self.G * self.L
        return out

    def elastic(self, Ein, out=None):
        Ein = f64(Ein)
        out = self._out(out, len(Ein))
        self._check(self.lib.ndppgpu_group_elastic(self.h, dp(Ein), len(Ein), dp(out)))
        return out

    def inelastic(self, Ein, out=None, nu_out=None):
        Ein = f64(Ein)
        out = self._out(out, len(Ein))
        nu = self._out(nu_out, len(Ein)) if self.params.nuscatter else None
        self._check(self.lib.ndppgpu_group_inelastic(self.h, dp(Ein), len(Ein), dp(out), dp(nu)))
        return out, nu

    # -- in pieces: grids and results stay on the devices -----------------------------------------------------------
    def set_grids(self, Ein_el=None, Ein_inel=None):
        a = f64(Ein_el) if Ein_el is not None else None
        b = f64(Ein_inel) if Ein_inel is not None else None
        # a negative count leaves that grid as it is
        self._check(self.lib.ndppgpu_group_set_grids(self.h, dp(a), len(a) if a is not None else -1, dp(b),
                                                     len(b) if b is not None else -1))
        if a is not None:
            self.n_el = len(a)
        if b is not None:
            self.n_inel = len(b)

    def integrate(self, what: int = 3):
        self._check(self.lib.ndppgpu_group_integrate(self.h, int(what)))

    def sync(self):
        self._check(self.lib.ndppgpu_group_sync(self.h))

    def join(self):
        """Device-side: every device's stream waits for the gathers enqueued so far (for CUDA-event timing)."""
        self._check(self.lib.ndppgpu_group_join(self.h))

    def fetch(self, el=True, inel=True, el_out=None, inel_out=None, nu_out=None):
        """Latest assembled matrices as host arrays (root; None elsewhere)."""
        e = self._out(el_out, self.n_el) if el and self.n_el else None
        i = self._out(inel_out, self.n_inel) if inel and self.n_inel else None
        n = self._out(nu_out, self.n_inel) if inel and self.n_inel and self.params.nuscatter else None
        self._check(self.lib.ndppgpu_group_fetch(self.h, dp(e), dp(i), dp(n)))
        return e, i, n

    def result_dev(self, matrix: int) -> int:
        return int(self.lib.ndppgpu_group_result_dev(self.h, int(matrix)) or 0)

    def clear(self):
        if self.h:
            self.lib.ndppgpu_group_nuclide_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.clear()
        except Exception:
            pass


# ---- library runs -------------------------------------------------------------------------------------------------------
def tile_bounds(n: int, tile: int, n_tiles: int) -> Tuple[int, int]:
    lo, hi = C.c_int(0), C.c_int(0)
    capi.load().ndppgpu_tile_bounds(int(n), int(tile), int(n_tiles), C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def plan_library(shapes, G: int, L: int, M: int, K: int, world: int, tile_rows: int = 1024, setup_cost: float = 5.0e10,
                 policy: str = "lpt"):
    """ndppgpu_plan_library on `library.NuclideShape`s.  Returns (items, imbalance): items are dicts with nuclide, matrix
    (0 elastic / 1 inelastic), tile, n_tiles, rank, rows (tile height on the shape's grid sizes), cost -- sorted by
    (rank, nuclide, matrix, tile).  Host code only: works without a GPU."""
    lib = capi.load()
    arr = (ShapeC * max(len(shapes), 1))()
    keep = []
    for k, s in enumerate(shapes):
        thr = f64(list(s.level_thresholds))
        keep.append(thr)
        arr[k] = ShapeC(int(s.index), int(s.n_el), int(s.n_inel), len(thr), int(s.cont_threshold is not None),
                        int(s.freegas_points), float(s.cont_threshold if s.cont_threshold is not None else 0.0),
                        float(s.e_lo), float(s.e_hi), dp(thr) if len(thr) else None)
    n = C.c_int(0)
    imb = C.c_double(0.0)
    pol = {"lpt": 0, "static": 1}[policy]
    check(lib.ndppgpu_plan_library(arr, len(shapes), G, L, M, K, tile_rows, world, float(setup_cost), pol, None, 0,
                                   C.byref(n), C.byref(imb)))
    items = (ItemC * max(n.value, 1))()
    check(lib.ndppgpu_plan_library(arr, len(shapes), G, L, M, K, tile_rows, world, float(setup_cost), pol, items, n.value,
                                   C.byref(n), C.byref(imb)))
    size = {s.index: (s.n_el, s.n_inel) for s in shapes}
    out = []
    for it in items[:n.value]:
        lo, hi = tile_bounds(size[it.nuclide][it.matrix], it.tile, it.n_tiles)
        out.append(dict(nuclide=it.nuclide, matrix=it.matrix, tile=it.tile, n_tiles=it.n_tiles, rank=it.rank,
                        rows=hi - lo, cost=it.cost))
    return out, imb.value


class LibraryRun:
    """ndppgpu_library_create / _run / _fetch.  `open_nuclide(index, ctx) -> (DeviceNuclide, Ein_el, Ein_inel)` builds
    nuclide `index` on the borrowed Context it is given (it is called from the worker thread of the device that needs
    it); the DeviceNuclide is freed by this class when the device is done with it."""

    def __init__(self, group: Group, G: int, L: int, nuscatter: bool, items: Sequence[dict]):
        self.group, self.lib, self.G, self.L = group, group.lib, G, L
        self.items = list(items)
        arr = (ItemC * max(len(self.items), 1))()
        for k, it in enumerate(self.items):
            arr[k] = ItemC(it["nuclide"], it["matrix"], it["tile"], it["n_tiles"], it["rank"], it["rows"], it["cost"])
        self.h = C.c_void_p()
        self._ctx0 = self.lib.ndppgpu_group_ctx(group.h, 0)
        check(self.lib.ndppgpu_library_create(group.h, G, L, int(bool(nuscatter)), arr, len(self.items), C.byref(self.h)),
              self._ctx0)

    def run(self, open_nuclide: Callable) -> dict:
        live = {}
        errors = []

        def _open(user, index, ctx_p, nuc_pp, el_pp, nel_p, in_pp, nin_p):
            try:
                dn, Eel, Ein = open_nuclide(int(index), Context(borrowed=ctx_p))
                Eel, Ein = f64(Eel), f64(Ein)
                live[(int(index), int(ctx_p))] = (dn, Eel, Ein)      # keeps the arrays alive until close
                nuc_pp[0] = dn.h.value
                el_pp[0] = Eel.ctypes.data_as(capi.c_dp)
                nel_p[0] = len(Eel)
                in_pp[0] = Ein.ctypes.data_as(capi.c_dp)
                nin_p[0] = len(Ein)
                return 0
            except Exception as e:   # an exception must not cross the C frames
                errors.append(e)
                return 1

        def _close(user, index, nuc_p):
            for key in [k for k in live if k[0] == int(index) and live[k][0].h.value == nuc_p]:
                live.pop(key)[0].clear()
            return 0

        rep = LibraryReportC()
        cb_open, cb_close = capi.OPEN_FN(_open), capi.CLOSE_FN(_close)
        rc = self.lib.ndppgpu_library_run(self.h, cb_open, cb_close, None, C.byref(rep))
        if errors:
            raise errors[0]
        check(rc, self._ctx0)
        return {k: getattr(rep, k) for k, _ in LibraryReportC._fields_ if k != "reserved"}

    def fetch(self, nuclide: int, matrix: int, Ein, e_top: float) -> np.ndarray:
        Ein = f64(Ein)
        out = np.empty((len(Ein), self.G, self.L))
        check(self.lib.ndppgpu_library_fetch(self.h, int(nuclide), int(matrix), dp(Ein), len(Ein), float(e_top), dp(out)),
              self._ctx0)
        return out

    def close(self):
        if self.h:
            self.lib.ndppgpu_library_free(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
