"""Host-side mirror of the reference's parsed ACE data model (the input of the hot path).

These plain containers carry exactly the fields of the reference's Fortran derived types that the
scattering-moment integrator reads -- `Tab1` (src/endf_header.F90:9-20), `DistAngle`
(src/ace_header.F90:14-24), `DistEnergy` (:31-44), `Reaction` (:50-69), `Nuclide` (:94-130),
`SAlphaBeta` (:201-235) -- with the same names and 1-based index conventions in the *values*
(`threshold`, `location`, locators inside `edist.data`).  ACE parsing itself (src/ace.F90) stays in
the reference's Fortran; these objects are what its ISO_C_BINDING shim flattens into the C-ABI of
include/ndppgpu.h, and what the synthetic generators in synth.py produce directly.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional

import numpy as np

# src/constants.F90:124-152
HISTOGRAM, LINEAR_LINEAR, LINEAR_LOG, LOG_LINEAR, LOG_LOG = 1, 2, 3, 4, 5
ANGLE_ISOTROPIC, ANGLE_32_EQUI, ANGLE_TABULAR = 1, 2, 3
SCATT_TYPE_LEGENDRE, SCATT_TYPE_TABULAR = 0, 1
SAB_SECONDARY_EQUAL, SAB_SECONDARY_SKEWED, SAB_SECONDARY_CONT = 0, 1, 2
SAB_ELASTIC_DISCRETE, SAB_ELASTIC_EXACT = 3, 4
ELASTIC, N_LEVEL, N_FISSION, N_NC = 2, 4, 18, 91
NU_NONE, NU_POLYNOMIAL, NU_TABULAR = 0, 1, 2   # src/constants.F90:187-189
K_BOLTZMANN = 8.617343e-11  # MeV/K (the reference reads kT from the ACE table)


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64).ravel())


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).ravel())


@dataclass
class Tab1:
    """ENDF TAB1 function (src/endf_header.F90:9-20)."""
    x: np.ndarray
    y: np.ndarray
    nbt: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))
    int: np.ndarray = field(default_factory=lambda: np.zeros(0, np.int32))

    def flatten(self) -> np.ndarray:
        """[NR, NBT(NR), INT(NR), NP, x(NP), y(NP)] -- the array form interpolate_tab1 reads
        (src/interpolation.F90:24-60)."""
        nr = len(self.nbt)
        return _f64(np.concatenate([[nr], self.nbt, self.int, [len(self.x)], self.x, self.y]))


@dataclass
class DistAngle:
    energy: np.ndarray    # incoming energies
    type: np.ndarray      # ANGLE_* per energy
    location: np.ndarray  # 0-based offset `lc` into data (data(lc+1) is the first value, 1-based)
    data: np.ndarray


@dataclass
class DistEnergy:
    law: int
    data: np.ndarray = field(default_factory=lambda: np.zeros(0))
    p_valid: Optional[Tab1] = None
    next: Optional["DistEnergy"] = None


@dataclass
class Reaction:
    MT: int
    Q_value: float = 0.0
    multiplicity: int = 1
    threshold: int = 1            # 1-based index into Nuclide.energy
    scatter_in_cm: bool = True
    sigma: np.ndarray = field(default_factory=lambda: np.zeros(0))
    adist: Optional[DistAngle] = None
    edist: Optional[DistEnergy] = None
    multiplicity_E: Optional[Tab1] = None

    @property
    def has_angle_dist(self) -> bool:
        return self.adist is not None

    @property
    def has_energy_dist(self) -> bool:
        return self.edist is not None


@dataclass
class Nuclide:
    awr: float
    kT: float
    energy: np.ndarray
    elastic: np.ndarray
    reactions: List[Reaction]
    freegas_cutoff: float = 0.0   # MeV (src/ndpp.F90:573-591 turns the xml value into MeV)
    name: str = "synthetic"
    # fission data read by calc_chi (src/ace_header.F90:113-130): nu_*_type are NU_NONE / NU_POLYNOMIAL / NU_TABULAR,
    # nu_*_data as the reference stores them ([NC, c...] or a flattened TAB1); fission = nuc % fission (the sum
    # of the fission cross sections on the nuclide grid; rebuilt from the reactions when None, src/ace.F90:823-828)
    fission: Optional[np.ndarray] = None
    nu_t_type: int = 0
    nu_t_data: Optional[np.ndarray] = None
    nu_d_type: int = 0
    nu_d_data: Optional[np.ndarray] = None
    nu_d_precursor_data: Optional[np.ndarray] = None   # per group: decay constant + TAB1 of its yield
    nu_d_edist: List[DistEnergy] = field(default_factory=list)

    @property
    def n_precursor(self) -> int:
        return len(self.nu_d_edist)

    @property
    def index_fission(self) -> List[int]:
        """0-based positions of the fission reactions (is_fission: MT 18, 19, 20, 21, 38)."""
        return [i for i, r in enumerate(self.reactions) if r.MT in (18, 19, 20, 21, 38)]

    @property
    def fissionable(self) -> bool:
        return bool(self.index_fission)


@dataclass
class DistEnergySab:
    e_out: np.ndarray
    e_out_pdf: np.ndarray
    mu: np.ndarray  # (n_mu, n_e_out) Fortran order == C array [n_e_out][n_mu]


@dataclass
class SAlphaBeta:
    awr: float
    kT: float
    threshold_inelastic: float
    inelastic_e_in: np.ndarray
    inelastic_sigma: np.ndarray
    secondary_mode: int
    n_inelastic_mu: int
    # discrete (equal / skewed) representation, C layout [n_e_in][n_e_out] and [n_e_in][n_e_out][n_mu]
    inelastic_e_out: Optional[np.ndarray] = None
    inelastic_mu: Optional[np.ndarray] = None
    # continuous representation
    inelastic_data: Optional[List[DistEnergySab]] = None
    # elastic
    threshold_elastic: float = 0.0
    elastic_mode: int = SAB_ELASTIC_DISCRETE
    elastic_e_in: Optional[np.ndarray] = None
    elastic_P: Optional[np.ndarray] = None
    elastic_mu: Optional[np.ndarray] = None  # C layout [n_e_in][n_mu]; None => coherent (n_elastic_mu = 0)
    name: str = "synthetic.sab"

    @property
    def n_inelastic_e_in(self) -> int:
        return len(self.inelastic_e_in)

    @property
    def n_inelastic_e_out(self) -> int:
        return 0 if self.inelastic_e_out is None else int(np.asarray(self.inelastic_e_out).shape[1])

    @property
    def n_elastic_e_in(self) -> int:
        return 0 if self.elastic_e_in is None else len(self.elastic_e_in)

    @property
    def n_elastic_mu(self) -> int:
        return 0 if self.elastic_mu is None else int(np.asarray(self.elastic_mu).shape[1])


@dataclass
class Params:
    """Run-time integration parameters, src/global.F90:28-59 with the defaults of
    src/constants.F90:69-100."""
    scatt_type: int = SCATT_TYPE_LEGENDRE
    order: int = 5
    mu_bins: int = 2001
    nuscatter: bool = False
    ne_per_grp: int = 20
    adaptive_mu_its: int = 15
    adaptive_eout_its: int = 15
    sab_threshold: float = 1.0e-6
    brent_mu_thresh: float = 1.0e-6
    adaptive_mu_tol: float = 1.0e-7
    adaptive_eout_tol: float = 1.0e-8


def iter_slots(nuc: Nuclide):
    """Yield (rxn_index, rxn, edist) in the order calc_scatt fills rxn_data(:)
    (src/scatt.F90:88-105): one slot per reaction, plus one per nested energy distribution."""
    for i, rxn in enumerate(nuc.reactions):
        yield i, rxn, rxn.edist
        ed = rxn.edist
        while ed is not None and ed.next is not None:
            ed = ed.next
            yield i, rxn, ed
