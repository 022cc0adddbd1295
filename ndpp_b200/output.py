"""NDPP library files: the wire format OpenMC reads.

Host-side mirror of the reference's writers, fed straight from the (thinned) moment arrays:

  group_index        <- src/ndpp.F90:649-683     energy-group locations in an E_in grid
  init_library       <- src/ndpp.F90:1246-1329   header: name(10), kT, NG, E_bins, scatt_type, scatt_order,
                                                 nuscatter, chi_present, mu_bins, thin_tol
  print_scatt        <- src/scatt.F90:821-997 (ASCII), :1139-1258 (BINARY, Fortran stream access)
                        per matrix: NE, Ein(:), grp_index(NG+1), per E_in: gmin, gmax (1-based; 0, 0 for an
                        all-zero column) + the L moments of every group of the window of positive P0
  read_library          restatement of the reference's reader src/utils/ndpp_data.py:141-245 (binary)

tests/test_output.py reads files written here with the reference's own reader (imported from
/root/reference in the build container) and compares every number.

  print_chi          <- src/chi.F90:165-337 (ASCII :195-236, BINARY :309-337): NE, number of precursor groups,
                        E_grid, chi_total(G, NE), chi_prompt(G, NE), chi_delay(G, NE, precursor), written after
                        the scattering data of a fissionable nuclide when integrate_chi is set (chi_present = 1,
                        src/ndpp.F90:1275-1279).  The reference's Python reader takes only NE values per chi
                        array (src/utils/ndpp_data.py:247-270), so chi records are checked with read_library here.
"""
from __future__ import annotations

import struct
from typing import Optional

import numpy as np

from .egrid import binary_search

ASCII, BINARY = "ascii", "binary"


def group_index(Ein, energy_bins) -> np.ndarray:
    """1-based group_index_* of the reference (src/ndpp.F90:649-683): for every group edge the location in
    Ein (binary_search), 1 below the grid, size(Ein) at or above its top; the last entry is size(Ein)."""
    Ein = np.asarray(Ein, dtype=np.float64)
    eb = np.asarray(energy_bins, dtype=np.float64)
    out = np.zeros(len(eb), dtype=np.int32)
    n = len(Ein)
    for g, e in enumerate(eb):
        if e < Ein[0]:
            out[g] = 1
        elif e >= Ein[-1]:
            out[g] = n
        else:
            out[g] = binary_search(Ein, e)       # 1-based, as src/search.F90
    out[-1] = n
    return out


def _fortran_e(v: float) -> str:
    """Fortran edit descriptor 1PE20.12."""
    s = f"{v:20.12E}"
    mant, exp = s.split("E")
    if len(exp) > 3:          # three-digit exponent: Fortran drops the letter, '1.000000000000+100'
        s = (mant.strip() + exp[0] + exp[1:].rjust(3, "0")).rjust(20)
    return s


def _ascii_array(f, arr, fmt):
    arr = list(arr)
    for i in range(0, len(arr), 4):
        f.write("".join(fmt(v) for v in arr[i:i + 4]).rstrip() + "\n")


def _windows(mat):
    """gmin, gmax (1-based, 0/0 for all-zero columns) from the P0 moments, src/scatt.F90:923-931."""
    pos = mat[:, :, 0] > 0.0
    any_pos = pos.any(axis=1)
    gmin = np.where(any_pos, pos.argmax(axis=1) + 1, 0)
    gmax = np.where(any_pos, pos.shape[1] - pos[:, ::-1].argmax(axis=1), 0)
    return gmin.astype(np.int32), gmax.astype(np.int32)


class LibraryWriter:
    """init_library + print_scatt for one nuclide (or S(a,b) table) file."""

    def __init__(self, filename, name: str, kT: float, energy_bins, scatt_type: int, scatt_order: int,
                 nuscatter: bool, mu_bins: int, thin_tol: float, lib_format: str = BINARY, sab: bool = False,
                 chi_present: bool = False):
        self.fmt = lib_format.lower()
        if self.fmt not in (ASCII, BINARY):
            raise ValueError("lib_format must be 'ascii' or 'binary' (HDF5 output is not built)")
        self.eb = np.asarray(energy_bins, dtype=np.float64)
        self.NG = len(self.eb) - 1
        self.nuscatter = bool(nuscatter) and not sab      # ndpp.F90:1272-1276
        name10 = (name + " " * 10)[:10]
        self.chi_present = bool(chi_present) and not sab  # integrate_chi .and. fissionable, ndpp.F90:1275-1279
        hdr_ints = (int(scatt_type), int(scatt_order), int(self.nuscatter), int(self.chi_present))
        if self.fmt == BINARY:
            self.f = open(filename, "wb")
            self.f.write(name10.encode("ascii"))
            self.f.write(struct.pack("=d", kT))
            self.f.write(struct.pack("=i", self.NG))
            self.f.write(self.eb.astype("=f8").tobytes())
            self.f.write(struct.pack("=4i", *hdr_ints))
            self.f.write(struct.pack("=i", int(mu_bins)))
            self.f.write(struct.pack("=d", float(thin_tol)))
        else:
            self.f = open(filename, "w")
            # '(A20,1PE20.12,I20,A20)': a character(10) name in an A20 field is right-justified
            self.f.write((f"{name10:>20s}" + _fortran_e(kT) + f"{self.NG:20d}").rstrip() + "\n")
            _ascii_array(self.f, self.eb, _fortran_e)
            self.f.write("".join(f"{v:20d}" for v in hdr_ints) + "\n")
            self.f.write(f"{int(mu_bins):20d}" + _fortran_e(float(thin_tol)) + "\n")

    def _matrix(self, mat):
        gmin, gmax = _windows(mat)
        for iE in range(mat.shape[0]):
            lo, hi = int(gmin[iE]), int(gmax[iE])
            if self.fmt == BINARY:
                self.f.write(struct.pack("=2i", lo, hi))
                if lo > 0:
                    self.f.write(np.ascontiguousarray(mat[iE, lo - 1:hi, :]).astype("=f8").tobytes())
            else:
                self.f.write(f"{lo:20d}{hi:20d}\n")
                if lo > 0:
                    _ascii_array(self.f, mat[iE, lo - 1:hi, :].ravel(), _fortran_e)

    def _grid(self, Ein):
        Ein = np.asarray(Ein, dtype=np.float64)
        gi = group_index(Ein, self.eb)
        if self.fmt == BINARY:
            self.f.write(struct.pack("=i", len(Ein)))
            self.f.write(Ein.astype("=f8").tobytes())
            self.f.write(gi.astype("=i4").tobytes())
        else:
            self.f.write(f"{len(Ein):20d}\n")
            _ascii_array(self.f, Ein, _fortran_e)
            _ascii_array(self.f, gi, lambda v: f"{int(v):20d}")

    def print_scatt(self, Ein_el, el_mat, Ein_inel=None, inel_mat=None, nuinel_mat=None):
        """Matrices are [NE][G][L] (the Fortran mat(L, G, NE))."""
        self._grid(Ein_el)
        self._matrix(np.asarray(el_mat))
        if Ein_inel is not None and len(Ein_inel) > 0:
            self._grid(Ein_inel)
            self._matrix(np.asarray(inel_mat))
            if self.nuscatter:
                if nuinel_mat is None:
                    raise ValueError("nuscatter is set but no nu-inelastic matrix was given")
                self._matrix(np.asarray(nuinel_mat))
        else:
            if self.fmt == BINARY:
                self.f.write(struct.pack("=i", 0))
            else:
                self.f.write(f"{0:20d}\n")

    def print_chi(self, E_grid, chi_t, chi_p, chi_d):
        """chi_t / chi_p are [NE][G] (the Fortran chi(G, NE)), chi_d is [precursor][NE][G]."""
        if not self.chi_present:
            raise ValueError("print_chi on a library whose header says chi_present = 0")
        E_grid = np.asarray(E_grid, dtype=np.float64)
        chi_t, chi_p, chi_d = (np.ascontiguousarray(a, dtype=np.float64) for a in (chi_t, chi_p, chi_d))
        n_prec = chi_d.shape[0] if chi_d.size else 0
        if self.fmt == BINARY:
            self.f.write(struct.pack("=2i", len(E_grid), n_prec))
            for a in (E_grid, chi_t, chi_p, chi_d):
                self.f.write(a.astype("=f8").tobytes())
        else:
            self.f.write(f"{len(E_grid):20d}{n_prec:20d}\n")
            _ascii_array(self.f, E_grid, _fortran_e)
            _ascii_array(self.f, chi_t.ravel(), _fortran_e)
            _ascii_array(self.f, chi_p.ravel(), _fortran_e)
            for k in range(n_prec):
                _ascii_array(self.f, chi_d[k].ravel(), _fortran_e)

    def close(self):
        self.f.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


def read_library(filename) -> dict:
    """Binary library -> dict, field for field as the reference's reader (src/utils/ndpp_data.py:141-245),
    with the windows expanded to dense [NE][G][L] arrays."""
    with open(filename, "rb") as f:
        def rd(fmt):
            return struct.unpack("=" + fmt, f.read(struct.calcsize("=" + fmt)))
        out = {"name": rd("10s")[0].decode("ascii"), "kT": rd("d")[0]}
        NG = rd("i")[0]
        out["NG"] = NG
        out["E_bins"] = np.array(rd(f"{NG + 1}d"))
        out["scatt_type"], out["scatt_order"], nus, chi = rd("4i")
        out["nuscatter"], out["chi_present"] = bool(nus), bool(chi)
        out["mu_bins"] = rd("i")[0]
        out["thin_tol"] = rd("d")[0]
        L = out["scatt_order"] + 1 if out["scatt_type"] == 0 else out["scatt_order"]

        def grid():
            NE = rd("i")[0]
            if NE == 0:
                return None, None
            return np.array(rd(f"{NE}d")), np.array(rd(f"{NG + 1}i"))

        def matrix(NE):
            m = np.zeros((NE, NG, L))
            for iE in range(NE):
                lo, hi = rd("2i")
                if lo > 0:
                    n = (hi - lo + 1) * L
                    m[iE, lo - 1:hi, :] = np.array(rd(f"{n}d")).reshape(hi - lo + 1, L)
            return m

        out["Ein_el"], out["grp_index_el"] = grid()
        out["elastic"] = matrix(len(out["Ein_el"]))
        out["Ein_inel"], out["grp_index_inel"] = grid()
        out["inelastic"] = out["nuinelastic"] = None
        if out["Ein_inel"] is not None:
            out["inelastic"] = matrix(len(out["Ein_inel"]))
            if out["nuscatter"]:
                out["nuinelastic"] = matrix(len(out["Ein_inel"]))
        if out["chi_present"]:
            NE, n_prec = rd("2i")
            out["Ein_chi"] = np.array(rd(f"{NE}d"))
            out["chi_total"] = np.array(rd(f"{NE * NG}d")).reshape(NE, NG)
            out["chi_prompt"] = np.array(rd(f"{NE * NG}d")).reshape(NE, NG)
            out["chi_delay"] = np.array(rd(f"{n_prec * NE * NG}d")).reshape(n_prec, NE, NG)
        out["trailing_bytes"] = len(f.read())
    return out
