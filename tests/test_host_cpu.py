"""CPU-side tests (no GPU): the C-ABI library loads and exports every declared symbol, the host grid
builders behave as the reference's, the oracle is self-consistent on the synthetic configurations,
and the multi-rank sharding logic works under gloo with world_size 2."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from ndpp_b200 import ace, capi, egrid, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from ndpp_b200 import build
    build.build()
    lib = capi.load()
    header = open(os.path.join(ROOT, "include", "ndppgpu.h")).read()
    declared = sorted(set(re.findall(r"\b(ndppgpu_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/ndppgpu.h but not exported"
    assert sorted(capi.EXPORTS) == declared
    assert lib.ndppgpu_abi_version() == 3


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(capi.NdppGpuError, match="no CUDA device"):
        capi.Context()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "ndpp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in src and "libndpp_oracle" not in src and "ndpp_oracle.h" not in src, f


def test_merge_and_search_follow_the_reference(oracle):
    import ctypes as C
    rng = np.random.default_rng(0)
    for _ in range(20):
        a = np.unique(rng.uniform(0, 1, rng.integers(2, 30)))
        b = np.unique(np.concatenate([rng.uniform(0, 1, rng.integers(2, 30)), a[:3]]))
        if rng.random() < 0.5:
            a = np.concatenate([[0.0], a])
        if rng.random() < 0.5:
            b = np.concatenate([[0.0], b])
        out = np.zeros(len(a) + len(b))
        n = oracle.lib().ref_merge(oracle.dp(a), len(a), oracle.dp(b), len(b), oracle.dp(out))
        assert np.array_equal(out[:n], egrid.merge(a, b))
        for v in (a[0], a[-1], a[len(a) // 2], 0.5 * (a[0] + a[-1])):
            assert egrid.binary_search(a, v) == oracle.lib().ref_binary_search(oracle.dp(a), len(a), C.c_double(v))


def test_create_ein_grid_shapes():
    from tests.util import small_heavy
    nuc = small_heavy()
    eb = synth.group_structure(70)
    el, inel = egrid.create_Ein_grid(nuc, eb)
    assert np.all(np.diff(el) > 0) and np.all(np.diff(inel) > 0)
    assert el[0] == egrid.MIN_EIN                       # the 0.0 group edge becomes MIN_EIN (array_merge.F90:49-53)
    assert el[-1] == 20.0 * (1.0 + float(np.float32(1e-3))) and inel[-1] == el[-1]   # ONE + 1.0E-3 in double (:438)
    thr = min(nuc.energy[r.threshold - 1] for r in nuc.reactions if r.MT != 2)
    assert inel[0] <= thr < inel[1] or inel[0] == thr
    assert len(inel) > len(el)                          # (INEL_EXTEND_PTS-1) points per level and group edge
    nuc3, eb3, _, _ = synth.c3_h1_freegas()
    el3, inel3 = egrid.create_Ein_grid(nuc3, eb3)
    assert inel3 is None                                # only elastic (scatt.F90:204)


def test_sab_egrid():
    sab = synth.c4_sab("skewed", n_ein=20)
    g = egrid.sab_egrid(sab, synth.group_structure(70))
    assert np.all(np.diff(g) >= 0) and g[-1] == sab.inelastic_e_in[-1]
    g0 = egrid.sab_egrid(sab, synth.group_structure(70), sab_epts_per_bin=0)
    assert len(g) == (len(g0) - 1) * egrid.EXTEND_PTS + len(g0)   # sab.F90:552


def test_oracle_properties_on_synthetic_configs(oracle):
    # free gas: normalised; S(a,b): normalised, last column copied; Law 44 CM: normalised x sigma
    nuc, eb, params, Ein = synth.c3_h1_freegas(n_ein=8)
    rn = oracle.RefNuclide(nuc, eb, params)
    rn.convert_distro()
    el = rn.elastic(Ein[[1, 6]])
    assert np.allclose(el[:, :, 0].sum(axis=1), 1.0, atol=1e-12)
    for mode in ("skewed", "cont"):
        sab = synth.c4_sab(mode, n_ein=12, n_eout=16 if mode == "skewed" else 80)
        E = np.geomspace(2e-11, 4e-6, 30)
        m = oracle.sab_calc(sab, eb, 5, E)
        assert np.allclose(m[:-1, :, 0].sum(axis=1), 1.0, atol=1e-12) and np.array_equal(m[-1], m[-2])


def test_freegas_explicit_stack_matches_recursion(oracle):
    """The device kernel unrolls the recursion onto explicit stacks; the same state machine in numpy
    is checked against the oracle's recursive form on a synthetic integrand."""
    def rec(f, a, b, eps, S, fa, fb, fc, bottom):
        c = 0.5 * (a + b); h = b - a; d = 0.5 * (a + c); e = 0.5 * (c + b)
        fd, fe = f(d), f(e)
        Sl = (h / 12.0) * (fa + 4.0 * fd + fc); Sr = (h / 12.0) * (fc + 4.0 * fe + fb); S2 = Sl + Sr
        if bottom <= 0 or abs(S2 - S) <= 15.0 * eps:
            return S2 + (S2 - S) / 15.0
        return rec(f, a, c, 0.5 * eps, Sl, fa, fc, fd, bottom - 1) + rec(f, c, b, 0.5 * eps, Sr, fc, fb, fe, bottom - 1)

    def stack(f, a0, b0, eps0, S0, fa0, fb0, fc0, bottom0):
        st = [dict(a=a0, b=b0, eps=eps0, S=S0, fa=fa0, fb=fb0, fc=fc0, bottom=bottom0, state=0, left=0.0)]
        have, val = False, 0.0
        while True:
            if have:
                if len(st) == 1 and st[0]["state"] == 0:
                    break
                if len(st) == 0:
                    break
                p = st[-1]
                if p["state"] == 1:
                    p["left"] = val; p["state"] = 2; have = False
                    st.append(dict(a=p["a"], b=p["b"], eps=p["eps"], S=p["S"], fa=p["fa"], fb=p["fb"], fc=p["fc"],
                                   bottom=p["bottom"], state=0, left=0.0))
                else:
                    val = p["left"] + val; st.pop()
                    if not st:
                        break
                continue
            f_ = st[-1]
            a, b = f_["a"], f_["b"]; c = 0.5 * (a + b); h = b - a; d = 0.5 * (a + c); e = 0.5 * (c + b)
            fd, fe = f(d), f(e)
            Sl = (h / 12.0) * (f_["fa"] + 4.0 * fd + f_["fc"]); Sr = (h / 12.0) * (f_["fc"] + 4.0 * fe + f_["fb"])
            S2 = Sl + Sr
            if f_["bottom"] <= 0 or abs(S2 - f_["S"]) <= 15.0 * f_["eps"]:
                val = S2 + (S2 - f_["S"]) / 15.0; have = True
                st.pop()
                if not st:
                    break
            else:
                fa_l, fc_l, fb_r, eps2, bot = f_["fa"], f_["fc"], f_["fb"], 0.5 * f_["eps"], f_["bottom"] - 1
                f_.update(a=c, b=b, eps=eps2, S=Sr, fa=fc_l, fb=fb_r, fc=fe, bottom=bot, state=1)
                st.append(dict(a=a, b=c, eps=eps2, S=Sl, fa=fa_l, fb=fc_l, fc=fd, bottom=bot, state=0, left=0.0))
        return val

    f = lambda x: np.exp(-40.0 * (x - 0.3) ** 2) * (1 + x)
    a, b = -1.0, 1.0
    fa, fb, fc = f(a), f(b), f(0.0)
    S = ((b - a) / 6.0) * (fa + 4 * fc + fb)
    assert rec(f, a, b, 1e-7, S, fa, fb, fc, 15) == stack(f, a, b, 1e-7, S, fa, fb, fc, 15)


GLOO_SCRIPT = r'''
# World-size-2 run of the host side of the one-process-per-GPU form (what an MPI driver does with MPI_Bcast /
# MPI_Allreduce around the collective ndppgpu_group_* / ndppgpu_library_* calls), over gloo on the CPU.
import os, sys
sys.path.insert(0, os.environ["NDPP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ndpp_b200 import group, library
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

# 1. the NCCL id is made on rank 0 (ncclGetUniqueId needs no GPU) and reaches every rank unchanged
nid = library.broadcast_id(dist)
ids = [None] * world
dist.all_gather_object(ids, nid)
assert len(nid) == 128 and any(nid) and all(x == nid for x in ids)

# 2. every rank derives the same plan from the shapes alone; each item has exactly one owner
shapes = [library.NuclideShape(i, 900 + 37 * i, 300 + 11 * i, (0.1 * (i + 1),), 1.5 if i % 2 else None, 1e-11, 20.0)
          for i in range(5)]
items, imb = library.plan(shapes, 4, 3, 101, 5, world, tile_rows=256)
plans = [None] * world
dist.all_gather_object(plans, [(it["rank"], it["nuclide"], it["matrix"], it["tile"], it["n_tiles"]) for it in items])
assert all(p == plans[0] for p in plans)
assert len({(n, m, t) for _, n, m, t, _ in plans[0]}) == len(plans[0])
assert {r for r, *_ in plans[0]} == set(range(world)) and 1.0 <= imb < 1.5

# 3. the true grid sizes are known only to the rank that parsed a nuclide (they differ from the planner's estimates):
#    after the exchange every rank holds all of them and the tile heights add up to the grids
own = library.owners(items)
true = {s.index: (s.n_el + 3 * s.index, s.n_inel + 5 + s.index) for s in shapes}
sizes = library.exchange_sizes(dist, len(shapes), {i: true[i] for i in own.get(rank, set())})
assert sizes == true
library.set_rows(items, sizes)
for s in shapes:
    for m in (0, 1):
        assert sum(it["rows"] for it in items if it["nuclide"] == s.index and it["matrix"] == m) == true[s.index][m]
rows = [None] * world
dist.all_gather_object(rows, [it["rows"] for it in items])
assert all(r == rows[0] for r in rows)

# 4. the cyclic deal of one nuclide's E_in grid (csrc/group.cuh: column i to device i mod N; the root puts column k of
#    device r back at i = r + k N) and the top-of-grid rule applied after the assembly, because the predecessor of a
#    column above the top group edge lives on another rank
NE, GL = 1003, 12
Ein = np.linspace(1.0, 3.0, NE); Ein[-1] = 3.003
mine = Ein[rank::world]
local = torch.tensor(np.outer(mine, np.arange(1, GL + 1)))       # stand-in for the moment columns of this rank
local[mine > 3.0] = float("nan")                                 # the copy rule must fill these
n_of = [(NE - r + world - 1) // world for r in range(world)]
assert local.shape[0] == n_of[rank]
pad = torch.zeros((max(n_of), GL), dtype=torch.float64); pad[:local.shape[0]] = local
stage = [torch.empty_like(pad) for _ in range(world)] if rank == 0 else None
dist.gather(pad, stage, dst=0)
if rank == 0:
    full = np.stack([stage[i % world][i // world].numpy() for i in range(NE)])
    for i in range(1, NE):
        if not Ein[i] <= 3.0:
            full[i] = full[i - 1]
    ref = np.outer(Ein, np.arange(1, GL + 1)); ref[-1] = ref[-2]
    assert np.array_equal(full, ref)
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_host_side_of_the_rank_form_world_size_2_gloo(tmp_path):
    script = tmp_path / "gloo_host.py"
    script.write_text(GLOO_SCRIPT)
    env = dict(os.environ, NDPP_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_library_plan_balances_better_than_static_blocks():
    '''ndppgpu_plan_library (host code of libndppgpu.so, no GPU needed): LPT over (nuclide, matrix, E_in tile) items vs
    the reference's contiguous nuclide blocks (src/ndpp.F90:941-948) on the C5 shapes: every item is planned exactly
    once, the plan is deterministic, and the modelled imbalance is far lower.'''
    from ndpp_b200 import group, library, synth
    specs = synth.c5_library(300)[:24]
    shapes = [synth.c5_shape(s) for s in specs]
    for world in (2, 8):
        lpt, imb = library.plan(shapes, 70, 6, 2001, 20, world)
        static, imb_s = library.plan(shapes, 70, 6, 2001, 20, world, policy="static")
        key = lambda it: (it["nuclide"], it["matrix"], it["tile"])
        assert sorted(map(key, lpt)) == sorted(map(key, static)) and len(set(map(key, lpt))) == len(lpt)
        again, _ = library.plan(list(reversed(shapes)), 70, 6, 2001, 20, world)
        assert [(it["rank"],) + key(it) for it in again] == [(it["rank"],) + key(it) for it in lpt]
        assert imb < 1.10 and imb <= imb_s
        # contiguous blocks of nuclides, first ranks take the remainder
        owner = {it["nuclide"]: it["rank"] for it in static}
        assert [owner[s.index] for s in shapes] == sorted(owner[s.index] for s in shapes)
        for it in lpt:
            n = shapes[it["nuclide"]].n_el if it["matrix"] == 0 else shapes[it["nuclide"]].n_inel
            assert it["rows"] == group.tile_bounds(n, it["tile"], it["n_tiles"])[1] - group.tile_bounds(n, it["tile"], it["n_tiles"])[0]
    assert imb_s > 1.3   # 8 ranks, 24 nuclides: the static blocks are far off
    lo = [group.tile_bounds(1003, t, 4) for t in range(4)]
    assert lo[0][0] == 0 and lo[-1][1] == 1003 and all(a[1] == b[0] for a, b in zip(lo, lo[1:]))
    # a nuclide whose inelastic slots have neither a level nor a continuum threshold still plans
    odd = [library.NuclideShape(0, 500, 200, (), None, 1e-11, 20.0)]
    it, _ = library.plan(odd, 70, 6, 2001, 20, 2)
    assert {x["matrix"] for x in it} == {0, 1}


def test_fused_closed_forms_are_current_and_bit_identical_on_the_host():
    """csrc/legendre_fused.inc (an off-by-default variant of the closed-form Legendre integrals with the exact
    power-of-two scalings folded into FMAs) is regenerated from the reference text in
    csrc/legendre.cuh and compared bit for bit with the oracle on the host."""
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("gen_legendre_fused", os.path.join(root, "scripts", "gen_legendre_fused.py"))
    g = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(g)
    text, marks = g.generate()
    assert open(g.OUT).read() == text, "legendre_fused.inc is stale"
    assert marks[7] < 330          # L = 8: 291 FP64 operations against 330 of the text as written
    rc, msg = g.check(300000)
    assert rc == 0 and msg.endswith(" 0 mismatches"), msg
