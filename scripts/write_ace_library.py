"""Write the benchmark nuclides as ACE type-1 files with the cross_sections.xml and ndpp.xml an unmodified
NDPP build reads, so that a maintainer with gfortran runs the reference on the very
nuclides the GPU path is measured on:

    python scripts/write_ace_library.py --config c2 --out /tmp/c2 && (cd /tmp/c2 && /path/to/ndpp)

The reference then builds its own E_in grids (create_Ein_grid); ndpp_b200.egrid restates that step.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from ndpp_b200 import acefile, synth  # noqa: E402


def write_ndpp_xml(path, e_bins, params, freegas_cutoff_kT, threads):
    with open(path, "w") as f:
        f.write('<?xml version="1.0" ?>\n<ndpp>\n  <scatt_type>legendre</scatt_type>\n')
        f.write(f"  <scatt_order>{params.order}</scatt_order>\n  <cross_sections>./cross_sections.xml</cross_sections>\n")
        f.write("  <energy_bins>" + " ".join(f"{e:.16E}" for e in e_bins) + "</energy_bins>\n")
        f.write(f"  <nuscatter>{'true' if params.nuscatter else 'false'}</nuscatter>\n  <integrate_chi>false</integrate_chi>\n")
        f.write(f"  <output_format>none</output_format>\n  <freegas_cutoff>{freegas_cutoff_kT}</freegas_cutoff>\n")
        f.write(f"  <mu_bins>{params.mu_bins}</mu_bins>\n  <print_tol>1.0E-10</print_tol>\n  <thinning_tol>0</thinning_tol>\n")
        f.write(f"  <threads>{threads}</threads>\n</ndpp>\n")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", choices=["c2", "c3", "c5"], default="c2")
    ap.add_argument("--out", required=True)
    ap.add_argument("--n-grid", type=int, default=20000)
    ap.add_argument("--n-nuclides", type=int, default=300)
    ap.add_argument("--threads", type=int, default=os.cpu_count() or 1)
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    tables = []
    if a.config == "c2":
        nuc, e_bins, params, _, _ = synth.c2_u238(n_grid=a.n_grid)
        tables.append(acefile.write_ace(nuc, os.path.join(a.out, "92238.70c.ace"), zaid=92238, name="92238.70c"))
        cutoff = 0.0
    elif a.config == "c3":
        for k, kT in enumerate((synth.KT_293K, 5.1704e-8, 1.0341e-7)):
            nuc, e_bins, params, _ = synth.c3_h1_freegas(kT=kT)
            tables.append(acefile.write_ace(nuc, os.path.join(a.out, f"1001.7{k}c.ace"), zaid=1001, name=f"1001.7{k}c"))
        cutoff = 400.0
    else:
        from ndpp_b200.ace import Params
        e_bins, params, cutoff = synth.group_structure(70), Params(order=5, mu_bins=2001), 0.0
        for spec in synth.c5_library(a.n_nuclides):
            nuc, _, _ = synth.c5_nuclide(spec)
            name = f"{1000 + spec[0]:d}.70c"
            tables.append(acefile.write_ace(nuc, os.path.join(a.out, name + ".ace"), zaid=1000 + spec[0], name=name))
    for t in tables:
        t["path"] = os.path.basename(t["path"])
    acefile.write_cross_sections_xml(tables, os.path.join(a.out, "cross_sections.xml"))
    write_ndpp_xml(os.path.join(a.out, "ndpp.xml"), e_bins, params, cutoff, a.threads)
    print(f"{len(tables)} ACE tables, cross_sections.xml and ndpp.xml written to {a.out}")


if __name__ == "__main__":
    main()
