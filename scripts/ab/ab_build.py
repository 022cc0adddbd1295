#!/usr/bin/env python
"""Builds named variants of libndppgpu.so for an A/B run on the GPU box.

    python scripts/ab/ab_build.py base= b4="-DF6_BLOCKS_PER_SM=4 -DF6_REG_PROD=40 -DF6_REG_CONS=88" plain="-DNDPP_FUSED_TABLELIN=0"
    gpurun -- 'VARIANTS="base b4 plain base" CHECK=b4 bash scripts/ab/ab_run.sh'

Each NAME=FLAGS pair is compiled with the extra nvcc flags into scripts/ab/libs/NAME.so (git-ignored, travels with gpurun);
scripts/ab/ab_run.sh copies the variants over ndpp_b200/csrc/libndppgpu.so one after the other, runs bench.py on each and the
bit-identity / parity subset of the GPU tests on $CHECK.  The default library is rebuilt at the end.  Registers and spill
of the dominant kernel are printed per variant (cuobjdump), so that a variant that cannot pay off is seen before any GPU
time is spent."""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ndpp_b200 import build as b  # noqa: E402

LIBS = os.path.join(ROOT, "scripts", "ab", "libs")


def main(argv):
    os.makedirs(LIBS, exist_ok=True)
    for arg in argv:
        name, _, flags = arg.partition("=")
        os.environ["NDPP_NVCC_EXTRA"] = flags
        b.build(force=True)
        dst = os.path.join(LIBS, name + ".so")
        shutil.copy(b.SO, dst)
        res = subprocess.run(["cuobjdump", "-res-usage", dst], capture_output=True, text=True).stdout.splitlines()
        use = [res[i + 1].strip() for i, l in enumerate(res) if "k_file6_cm_wsILi8E" in l and i + 1 < len(res)]
        print(f"{name}: {flags or '(default flags)'}\n    k_file6_cm_ws<8>: {use[0] if use else '?'}", flush=True)
    os.environ["NDPP_NVCC_EXTRA"] = ""
    b.build(force=True)
    b.build_tool()


if __name__ == "__main__":
    main(sys.argv[1:])
