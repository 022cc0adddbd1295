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
    assert lib.ndppgpu_abi_version() == 1


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
    assert el[-1] == np.float64(20.0) * (1.0 + np.float32(1e-3)) and inel[-1] == el[-1]
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
import os, sys
sys.path.insert(0, os.environ["NDPP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ndpp_b200 import parallel
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
NE, GL = 1003, 12
Ein = np.linspace(1.0, 3.0, NE); Ein[-1] = 3.003
lo, hi = parallel.shard_range(NE, rank, world)
assert parallel.shard_range(NE, 0, world)[0] == 0 and parallel.shard_range(NE, world - 1, world)[1] == NE
local = torch.tensor(np.outer(Ein[lo:hi], np.arange(1, GL + 1)))      # stand-in for the moment slab
local[Ein[lo:hi] > 3.0] = float("nan")                               # columns the copy rule must fill
full = parallel.gather_columns(local, NE, dst=0)
if rank == 0:
    parallel.copy_top_columns(full, torch.tensor(Ein), 3.0)
    ref = np.outer(Ein, np.arange(1, GL + 1)); ref[-1] = ref[-2]
    assert np.array_equal(full.numpy(), ref)
    print("GLOO_OK")
dist.destroy_process_group()
'''


def test_shard_and_gather_world_size_2_gloo(tmp_path):
    script = tmp_path / "gloo_shard.py"
    script.write_text(GLOO_SCRIPT)
    env = dict(os.environ, NDPP_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env,
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_OK" in out.stdout


def test_library_plan_balances_better_than_static_blocks():
    """LPT over (nuclide, matrix, E_in tile) items vs the reference's contiguous nuclide blocks
    (src/ndpp.F90:941-948) on the C5 shapes: every item is planned exactly once, the plan is
    deterministic, and the modelled imbalance is far lower."""
    from ndpp_b200 import library, synth
    specs = synth.c5_library(300)[:24]
    shapes = [synth.c5_shape(s) for s in specs]
    items = library.make_items(shapes, 70, 6, 2001, 20)
    for world in (2, 8):
        lpt, static = library.plan_lpt(items, world), library.plan_static_blocks(items, shapes, world)
        assert sorted((i.nuclide, i.matrix, i.tile) for p in lpt for i in p) == \
               sorted((i.nuclide, i.matrix, i.tile) for i in items)
        assert lpt == library.plan_lpt(list(reversed(items)), world)
        assert library.imbalance(lpt) < 1.10 and library.imbalance(lpt) <= library.imbalance(static)
    assert library.imbalance(static) > 1.3   # 8 ranks, 24 nuclides: the static blocks are far off
    lo = [library.tile_bounds(1003, t, 4) for t in range(4)]
    assert lo[0][0] == 0 and lo[-1][1] == 1003 and all(a[1] == b[0] for a, b in zip(lo, lo[1:]))


GLOO_LIBRARY = r'''
import os, sys
sys.path.insert(0, os.environ["NDPP_ROOT"])
import numpy as np, torch, torch.distributed as dist
from ndpp_b200 import library
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
GL = 6
shapes = [library.NuclideShape(i, 900 + 37 * i, 300 + 11 * i, (0.1 * (i + 1),), 1.5 if i % 2 else None, 1e-11, 20.0)
          for i in range(5)]
items = library.make_items(shapes, 4, 3, 101, 5, tile_rows=256)
plan = library.plan_lpt(items, world)
grids = {s.index: (np.geomspace(1e-11, 20.003, s.n_el), np.geomspace(0.2, 20.003, s.n_inel + 3)) for s in shapes}
opened = []
def open_nuclide(i):
    opened.append(i); return i
def integrate(i, it):               # stand-in for the device integration: rows identify (nuclide, E_in)
    E = grids[i][0 if it.matrix == "el" else 1]
    lo, hi = library.tile_bounds(len(E), it.tile, it.n_tiles)
    out = torch.tensor(np.outer(E[lo:hi], np.arange(1, GL + 1)) + 1000.0 * i)
    out[E[lo:hi] > 20.0] = float("nan")      # the copy rule must fill these from the predecessor
    return out
got = library.run_plan(plan[rank], plan, open_nuclide, integrate, lambda h: None, GL, torch.device("cpu"))
assert len(opened) == len(set(opened)), "a nuclide was opened twice on one rank"
opened.clear()
# the in-place form (heights from the host-side grids, one result buffer per rank) assembles the same library
def rows_of(it):
    lo, hi = library.tile_bounds(len(grids[it.nuclide][0 if it.matrix == "el" else 1]), it.tile, it.n_tiles)
    return hi - lo
def integrate_into(i, it, out):
    out.copy_(integrate(i, it))
got2 = library.run_plan(plan[rank], plan, open_nuclide, integrate_into, lambda h: None, GL, torch.device("cpu"),
                        rows_of=rows_of)
if rank == 0:
    assert set(got2) == set(got)
    for key in got:
        a = torch.cat([p[2] for p in sorted(got[key], key=lambda p: p[0])])
        b = torch.cat([p[2] for p in sorted(got2[key], key=lambda p: p[0])])
        assert torch.equal(torch.nan_to_num(a, nan=-1.0), torch.nan_to_num(b, nan=-1.0)), key
if rank == 0:
    assert set(got) == {(s.index, m) for s in shapes for m in ("el", "inel")}
    for (i, m), pieces in got.items():
        E = grids[i][0 if m == "el" else 1]
        mat = library.assemble(pieces, E, 20.0).numpy()
        ref = np.outer(E, np.arange(1, GL + 1)) + 1000.0 * i
        ref[-1] = ref[-2]
        assert np.array_equal(mat, ref), (i, m)
    print("GLOO_LIB_OK")
dist.destroy_process_group()
'''


def test_library_run_plan_world_size_2_gloo(tmp_path):
    script = tmp_path / "gloo_lib.py"
    script.write_text(GLOO_LIBRARY)
    env = dict(os.environ, NDPP_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29534", str(script)], env=env,
                         capture_output=True, text=True, timeout=240)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "GLOO_LIB_OK" in out.stdout


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
