#!/usr/bin/env python
"""Turns ncu output into the small text summaries kept under profiles/.

  launches  <launches.csv> <out.md> [title]    per-kernel launch counts, total time and share of the
                                               integrator kernels (from `ncu --metrics gpu__time_duration.sum --csv`)
  full      <report.ncu-rep> <out.md> [title]  key counters of every kernel in a `--set full` report
                                               (needs `ncu` on PATH to read the report)
"""
import collections
import csv
import subprocess
import sys

SKIP = ("k_fp64_peak", "at::native", "vectorized_elementwise", "k_test_")


def launches(path, out, title):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows:
        name = r[ik].split("(")[0].replace("void ", "")
        ns = float(r[iv].replace(",", "")) * {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(r[iu], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += ns
    core = {k: v for k, v in agg.items() if not any(s in k for s in SKIP)}
    tot = sum(v[1] for v in core.values())
    with open(out, "w") as f:
        f.write(f"# {title}\n\nTimes are cold-cache and serialised by ncu; compare shares, not absolutes.\n"
                "Kernels outside the step (FP64-peak micro-benchmark, torch fill) are listed but left out of the share.\n\n"
                "| kernel | launches | total ms | share of the integrator kernels |\n|---|---|---|---|\n")
        for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            share = f"{v[1] / tot:.3f}" if k in core else "-"
            f.write(f"| `{k}` | {v[0]} | {v[1] / 1e6:.3f} | {share} |\n")
        f.write(f"\nIntegrator kernels total: {tot / 1e6:.3f} ms.\n")


KEYS = [("gpu__time_duration.sum", "duration"), ("launch__registers_per_thread", "registers/thread"),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
        ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "FP64 pipe active %"),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
        ("smsp__inst_executed.sum", "warp instructions"),
        ("dram__bytes_read.sum", "DRAM read"), ("dram__bytes_write.sum", "DRAM write"),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %"), ("lts__t_sector_hit_rate.pct", "L2 hit %"),
        ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %")]


def full(rep, out, title):
    if rep.endswith(".csv"):     # `ncu -i report --page raw --csv` exported on the GPU box (the reports are too large to bring back)
        txt = open(rep).read()
    else:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, rows = rows[0], rows[1], rows[2:]
    with open(out, "w") as f:
        f.write(f"# {title}\n\nSource: `{rep.split('/')[-1]}` (`ncu --set full --clock-control none`; scripts/profile_r2.sh).\n")
        for r in rows:
            d = dict(zip(hdr, r))
            u = dict(zip(hdr, units))
            f.write(f"\n## `{d['Kernel Name'][:110]}`\n\ngrid {d['Grid Size']} x block {d['Block Size']}\n\n| counter | value |\n|---|---|\n")
            for k, label in KEYS:
                if k in d:
                    f.write(f"| {label} (`{k}`) | {d[k]} {u[k]} |\n")
            f.write("\nWarp stall reasons (warps per issue-active cycle):\n\n| reason | value |\n|---|---|\n")
            st = [(k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), float(d[k]))
                  for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
            for k, v in sorted(st, key=lambda kv: -kv[1])[:8]:
                f.write(f"| {k} | {v:.3f} |\n")


if __name__ == "__main__":
    mode, src, dst = sys.argv[1:4]
    title = sys.argv[4] if len(sys.argv) > 4 else src
    (launches if mode == "launches" else full)(src, dst, title)
