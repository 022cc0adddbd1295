"""ACE type-1 (ASCII) files of continuous-energy neutron tables.

`write_ace` lays an `ace.Nuclide` out as the NXS / JXS / XSS arrays that the reference's reader parses
(src/ace.F90:227-403 header and XSS, :410-487 ESZ, :685-861 MTR/LQR/TYR/LSIG/SIG, :868-956 LAND/AND,
:963-1277 LDLW/DLW with laws 3, 4, 9, 44, 61 and nested laws) and `write_cross_sections_xml` writes the
matching listing (src/ndpp.F90:1130-1236), so that an unmodified NDPP build can consume the very nuclides
the GPU path is measured on.  `read_ace` restates the reference's parse of those blocks (including the
locator re-basing of read_angular_dist :945-954 and length_energy_dist :1118-1246) and returns the same
`ace.Nuclide` model; it is what the round-trip tests and the benchmarks use, since no Fortran compiler
exists in the build image.

Numbers are written with the customary 1PE20.12, so a nuclide read back from a file carries 13 significant
digits; reading, writing and reading again is the identity.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .ace import ANGLE_32_EQUI, ANGLE_ISOTROPIC, ANGLE_TABULAR, ELASTIC, DistAngle, DistEnergy, Nuclide, Reaction, Tab1

_LOCATOR_LAWS = (4, 44, 61)


def _tab1_block(t: Optional[Tab1], x_default: Tuple[float, float]) -> List[float]:
    """[NR, NBT, INT, NE, x, y] of a law-validity / yield TAB1 (probability 1 when absent)."""
    if t is None:
        return [0.0, 2.0, x_default[0], x_default[1], 1.0, 1.0]
    nr = len(t.nbt)
    return [float(nr)] + [float(v) for v in t.nbt] + [float(v) for v in t.int] + [float(len(t.x))] + \
        [float(v) for v in t.x] + [float(v) for v in t.y]


def _law_data_to_file(law: int, data: np.ndarray, shift: int) -> List[float]:
    """edist%data with its row locators moved from data-relative (as stored after the reference's
    length_energy_dist re-basing) back to DLW-relative: L_file = L_data + LOCC + lid (`shift`)."""
    d = [float(v) for v in np.asarray(data, dtype=np.float64)]
    if law not in _LOCATOR_LAWS:
        return d
    NR = int(d[0])
    NE = int(d[1 + 2 * NR])
    loc0 = 2 + 2 * NR + NE
    for i in range(NE):
        lc = int(d[loc0 + i])
        if law == 61:
            NP = int(d[lc + 1])
            for j in range(NP):
                k = lc + 2 + 3 * NP + j
                if d[k] == 0.0:
                    raise ValueError("law 61: an outgoing energy without an angular table (LC = 0) cannot be written: the "
                                     "reference's length_energy_dist expects one table per outgoing energy "
                                     "(src/ace.F90:1233-1238)")
                d[k] = d[k] + shift
        d[loc0 + i] = float(lc + shift)
    return d


def _law_data_from_file(law: int, d: List[float], shift: int) -> Tuple[np.ndarray, int]:
    """Inverse of _law_data_to_file on the XSS slice starting at the law data; also returns its length.
    Follows length_energy_dist (src/ace.F90:1080-1277) for the laws the path accepts: the rows are walked
    in order and assumed back to back, a locator shared with a later row is re-based without adding length
    (:1131-1141), and for law 61 one angular table is expected behind every outgoing energy (:1233-1238)."""
    if law in (2, 3, 66):
        return np.array(d[:2]), 2
    if law in (7, 9):
        NR = int(d[0]); NE = int(d[1 + 2 * NR])
        n = 3 + 2 * NR + 2 * NE
        return np.array(d[:n]), n
    if law not in _LOCATOR_LAWS:
        raise ValueError(f"energy law {law} is not handled by this reader")
    NR = int(d[0]); NE = int(d[1 + 2 * NR])
    loc0 = 2 + 2 * NR + NE
    out = list(d)
    L = [int(v) for v in out[loc0:loc0 + NE]]
    length = 2 + 2 * NR + 2 * NE
    for i in range(NE):
        if i < NE - 1 and L[i] in L[i + 1:]:
            out[loc0 + i] -= shift
            continue
        NP = int(out[length + 1])
        if law == 4:
            length += 2 + 3 * NP
        elif law == 44:
            length += 2 + 5 * NP
        else:
            for j in range(NP):
                k = length + 2 + 3 * NP + j
                if out[k] != 0.0:
                    out[k] -= shift
            length += 2 + 4 * NP
            for j in range(NP):
                length += 2 + 3 * int(out[length + 1])
        out[loc0 + i] -= shift
    return np.array(out[:length]), length


def pack_adist(ad: DistAngle) -> DistAngle:
    """The angular tables laid out back to back in energy order, as read_angular_dist sizes the block
    (33 values per equiprobable table, 2 + 3 NP per tabular one); locations are recomputed."""
    data: List[float] = []
    loc = []
    for t, lc in zip(ad.type, ad.location):
        lc = int(lc)
        if t == ANGLE_ISOTROPIC:
            loc.append(0)
            continue
        loc.append(len(data))
        n = 33 if t == ANGLE_32_EQUI else 2 + 3 * int(ad.data[lc + 1])
        data += [float(v) for v in ad.data[lc:lc + n]]
    return DistAngle(energy=np.asarray(ad.energy, dtype=np.float64), type=np.asarray(ad.type, np.int32),
                     location=np.asarray(loc, np.int32), data=np.asarray(data, dtype=np.float64))


def build_xss(nuc: Nuclide, zaid: int = 0):
    """NXS(16), JXS(32), XSS of a nuclide.  reactions[0] must be elastic; the other reactions are written in
    their order, those with secondary neutrons (an energy law) first as the format requires."""
    rx = nuc.reactions
    if not rx or rx[0].MT != ELASTIC:
        raise ValueError("reactions[0] must be the elastic reaction (MT 2)")
    others = rx[1:]
    with_n = [r for r in others if r.edist is not None]
    without = [r for r in others if r.edist is None]
    for r in without:
        if r.adist is not None:
            raise ValueError(f"MT {r.MT}: an angular distribution without an energy law cannot be written to an ACE "
                             "table (every reaction with secondary neutrons carries a law; use law 3 for levels)")
    ordered = with_n + without
    if [id(r) for r in ordered] != [id(r) for r in others]:
        raise ValueError("reactions with secondary neutrons must precede the others (ACE NXS(5) ordering)")
    NE, NTR, NRn = len(nuc.energy), len(ordered), len(with_n)
    e_lo, e_hi = float(nuc.energy[0]), float(nuc.energy[-1])

    xss: List[float] = []
    jxs = [0] * 32
    # ESZ: energy, total, absorption, elastic, heating (total / absorption are rebuilt by the reader)
    jxs[0] = 1
    xss += [float(v) for v in nuc.energy] + [0.0] * (2 * NE) + [float(v) for v in nuc.elastic] + [0.0] * NE
    # MTR, LQR, TYR
    jxs[2] = len(xss) + 1
    xss += [float(r.MT) for r in ordered]
    jxs[3] = len(xss) + 1
    xss += [float(r.Q_value) for r in ordered]
    jxs[4] = len(xss) + 1
    tyr_at = len(xss)
    xss += [0.0] * NTR            # filled once the yield tables have their DLW offsets
    # LSIG, SIG
    jxs[5] = len(xss) + 1
    lsig_at = len(xss)
    xss += [0.0] * NTR
    jxs[6] = len(xss) + 1
    sig0 = len(xss)
    for i, r in enumerate(ordered):
        xss[lsig_at + i] = float(len(xss) - sig0 + 1)
        xss += [float(r.threshold), float(len(r.sigma))] + [float(v) for v in r.sigma]
    # LAND, AND
    jxs[7] = len(xss) + 1
    land_at = len(xss)
    xss += [0.0] * (NRn + 1)
    jxs[8] = len(xss) + 1
    and0 = len(xss)
    for i, r in enumerate([rx[0]] + with_n):
        if r.adist is None:
            xss[land_at + i] = -1.0 if (r.edist is not None and r.edist.law == 44) else 0.0
            continue
        ad = pack_adist(r.adist)
        LOCB = len(xss) - and0 + 1
        xss[land_at + i] = float(LOCB)
        n = len(ad.energy)
        base = LOCB + 2 * n + 1       # read_angular_dist: location = |LC| - (LOCB + 2 NE + 1)
        lcs = []
        for t, lc in zip(ad.type, ad.location):
            if t == ANGLE_ISOTROPIC:
                lcs.append(0.0)
            elif t == ANGLE_32_EQUI:
                lcs.append(float(int(lc) + base))
            elif t == ANGLE_TABULAR:
                lcs.append(-float(int(lc) + base))
            else:
                raise ValueError(f"unknown angular distribution type {t}")
        xss += [float(n)] + [float(v) for v in ad.energy] + lcs + [float(v) for v in ad.data]
    # LDLW, DLW
    jxs[9] = len(xss) + 1
    ldlw_at = len(xss)
    xss += [0.0] * NRn
    jxs[10] = len(xss) + 1
    dlw0 = len(xss)
    for i, r in enumerate(with_n):
        ed = r.edist
        first = True
        while ed is not None:
            LOCC = len(xss) - dlw0 + 1
            if first:
                xss[ldlw_at + i] = float(LOCC)
                first = False
            pv = _tab1_block(ed.p_valid, (e_lo if r.threshold <= 1 else float(nuc.energy[r.threshold - 1]), e_hi))
            lid = 3 + len(pv)          # LNW, LAW, IDAT + the TAB1 (= 5 + 2 (NR + NE) of the reference)
            data = _law_data_to_file(ed.law, ed.data, LOCC + lid)
            IDAT = LOCC + lid          # 1-based, relative to JXS(11)
            nxt = LOCC + lid + len(data) if ed.next is not None else 0
            xss += [float(nxt), float(ed.law), float(IDAT)] + pv + data
            ed = ed.next
    # energy-dependent yields (TYR = +-(100 + offset into DLW))
    for i, r in enumerate(ordered):
        mult = int(r.multiplicity)
        if r.multiplicity_E is not None:
            off = len(xss) - dlw0 + 1
            xss += _tab1_block(r.multiplicity_E, (e_lo, e_hi))
            mult = 100 + off
        xss[tyr_at + i] = float(-mult if r.scatter_in_cm else mult)
    jxs[21] = len(xss)             # END
    nxs = [0] * 16
    nxs[0], nxs[1], nxs[2], nxs[3], nxs[4] = len(xss), int(zaid), NE, NTR, NRn
    return nxs, jxs, xss


def _e20(v: float) -> str:
    s = f"{v:20.12E}"
    mant, exp = s.split("E")
    if len(exp) > 3:
        s = (mant.strip() + "E" + exp[0] + exp[-3:])[-20:].rjust(20)   # keep the letter: G20.0 reads either form
    return s


def write_ace(nuc: Nuclide, path: str, zaid: int = 0, name: Optional[str] = None, date: str = "18/10/26",
              comment: str = "synthetic ACE table written by ndpp_b200.acefile", mat: str = "   mat   0") -> dict:
    """Write one type-1 table; returns the cross_sections.xml attributes of the table."""
    nxs, jxs, xss = build_xss(nuc, zaid)
    name10 = ((name or nuc.name) + " " * 10)[:10]
    with open(path, "w") as f:
        # (A10,2G12.0,1X,A10)
        f.write(f"{name10}{nuc.awr:12.6f}{nuc.kT:12.4E} {date:<10.10s}\n")
        f.write(f"{comment:<70.70s}{mat:<10.10s}\n")
        for _ in range(4):
            f.write("".join(f"{0:7d}{0.0:11.0f}" for _ in range(4)) + "\n")
        for row in range(2):
            f.write("".join(f"{v:9d}" for v in nxs[8 * row:8 * row + 8]) + "\n")
        for row in range(4):
            f.write("".join(f"{v:9d}" for v in jxs[8 * row:8 * row + 8]) + "\n")
        for i in range(0, len(xss), 4):
            f.write("".join(_e20(v) for v in xss[i:i + 4]) + "\n")
    return {"name": name10.strip(), "alias": name10.strip(), "zaid": int(zaid), "type": "neutron", "awr": nuc.awr,
            "temperature": nuc.kT, "path": path, "location": 1, "filetype": "ascii"}


def write_cross_sections_xml(tables: List[dict], path: str, directory: str = ""):
    """The listing the reference's read_cross_sections_xml parses (src/ndpp.F90:1130-1236)."""
    with open(path, "w") as f:
        f.write('<?xml version="1.0" ?>\n<cross_sections>\n')
        if directory:
            f.write(f"  <directory>{directory}</directory>\n")
        f.write("  <filetype>ascii</filetype>\n")
        for t in tables:
            f.write(f'  <ace_table alias="{t["alias"]}" awr="{t["awr"]:.6f}" location="{t["location"]}" name="{t["name"]}" '
                    f'path="{t["path"]}" temperature="{t["temperature"]:.6e}" zaid="{t["zaid"]}"/>\n')
        f.write("</cross_sections>\n")


def read_ace(path: str, location: int = 1) -> Nuclide:
    """Parse one ASCII table as read_ace_table does (neutron tables; blocks listed in the module docstring)."""
    with open(path) as f:
        lines = f.read().split("\n")
    ln = lines[location - 1:]
    h = ln[0]
    name, awr, kT = h[0:10], float(h[10:22]), float(h[22:34])
    ints: List[int] = []
    for row in ln[6:12]:
        ints += [int(row[9 * k:9 * k + 9]) for k in range(8)]
    nxs, jxs = ints[:16], ints[16:48]
    n = nxs[0]
    xss: List[float] = []
    row = 12
    while len(xss) < n:
        s = ln[row]
        xss += [float(s[20 * k:20 * k + 20]) for k in range(len(s) // 20)]
        row += 1
    xss = xss[:n]
    X = lambda i: xss[i - 1]                      # 1-based, as the Fortran indexes XSS
    NE, NTR, NRn = nxs[2], nxs[3], nxs[4]
    energy = np.array(xss[0:NE])
    elastic = np.array(xss[3 * NE:4 * NE])
    LMT, JXS4, JXS5, LXS, JXS7, JXS8, JXS9, LED, LDIS = jxs[2], jxs[3], jxs[4], jxs[5], jxs[6], jxs[7], jxs[8], jxs[9], jxs[10]

    def tab1_at(i):            # XSS index of NR -> (Tab1, next index)
        NR = int(X(i))
        nbt = [int(X(i + 1 + k)) for k in range(NR)]
        itp = [int(X(i + 1 + NR + k)) for k in range(NR)]
        NP = int(X(i + 1 + 2 * NR))
        x = [X(i + 2 + 2 * NR + k) for k in range(NP)]
        y = [X(i + 2 + 2 * NR + NP + k) for k in range(NP)]
        return Tab1(x=np.array(x), y=np.array(y), nbt=np.array(nbt, np.int32), int=np.array(itp, np.int32)), \
            i + 2 + 2 * NR + 2 * NP

    rxns = [Reaction(MT=ELASTIC, Q_value=0.0, multiplicity=1, threshold=1, scatter_in_cm=True, sigma=np.zeros(0))]
    for i in range(1, NTR + 1):
        ty = int(round(X(JXS5 + i - 1)))
        r = Reaction(MT=int(X(LMT + i - 1)), Q_value=X(JXS4 + i - 1), multiplicity=abs(ty), scatter_in_cm=ty < 0)
        if r.multiplicity > 100:
            r.multiplicity_E, _ = tab1_at(LDIS + r.multiplicity - 101)
        LOCA = int(X(LXS + i - 1))
        r.threshold = int(X(JXS7 + LOCA - 1))
        ne = int(X(JXS7 + LOCA))
        r.sigma = np.array(xss[JXS7 + LOCA:JXS7 + LOCA + ne])
        rxns.append(r)
    # angular distributions (read_angular_dist)
    for i in range(1, NRn + 2):
        LOCB = int(X(JXS8 + i - 1))
        if LOCB <= 0:
            continue
        ne = int(X(JXS9 + LOCB - 1))
        e = xss[JXS9 + LOCB - 1:JXS9 + LOCB - 1 + ne]
        lc = [int(v) for v in xss[JXS9 + LOCB - 1 + ne:JXS9 + LOCB - 1 + 2 * ne]]
        types, length = [], 0
        for c in lc:
            if c == 0:
                types.append(ANGLE_ISOTROPIC)
            elif c > 0:
                types.append(ANGLE_32_EQUI); length += 33
            else:
                types.append(ANGLE_TABULAR); length += 2 + 3 * int(X(JXS9 + abs(c)))
        start = JXS9 + LOCB + 2 * ne          # XSS_index of the data (1-based)
        data = np.array(xss[start - 1:start - 1 + length])
        base = LOCB + 2 * ne + 1
        loc = [0 if c == 0 else abs(c) - base for c in lc]
        rxns[i - 1].adist = DistAngle(energy=np.array(e), type=np.array(types, np.int32), location=np.array(loc, np.int32),
                                      data=data)
    # energy distributions (read_energy_dist / get_energy_dist)
    def get_edist(loc_law):
        LNW, LAW, IDAT = int(X(LDIS + loc_law - 1)), int(X(LDIS + loc_law)), int(X(LDIS + loc_law + 1))
        pv, after = tab1_at(LDIS + loc_law + 2)
        lid = 3 + (after - (LDIS + loc_law + 2))
        lc = LDIS + IDAT - 2                  # data starts at XSS(lc + 1)
        data, _ = _law_data_from_file(LAW, xss[lc:], loc_law + lid)
        ed = DistEnergy(law=LAW, data=data, p_valid=pv)
        if LNW > 0:
            ed.next = get_edist(LNW)
        return ed
    for i in range(1, NRn + 1):
        rxns[i].edist = get_edist(int(X(LED + i - 1)))
    return Nuclide(awr=awr, kT=kT, energy=energy, elastic=elastic, reactions=rxns, name=name.strip())
