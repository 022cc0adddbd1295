#!/usr/bin/env python
"""C5: the synthetic library (examples/ndpp.xml shape, P5, 70 groups) over the GPUs of one box, through the C-ABI's
multi-device entry points (ndppgpu_group_* / ndppgpu_library_*, csrc/group.cuh).

  python scripts/run_library.py --devices 8            one process, 8 GPUs (host threads + ncclCommInitAll in the library)
  torchrun --nproc-per-node 8 scripts/run_library.py   one process per GPU (ncclCommInitRank; id and grid sizes are
                                                       exchanged over torch.distributed as an MPI driver would over MPI)

Prints one JSON line on the root: whole-job moment evaluations per second (host wall clock around ndppgpu_library_run,
max over ranks), the modelled imbalance of the LPT plan next to the reference's static nuclide blocks
(src/ndpp.F90:941-948), and a parity check of sampled nuclides against the CPU oracle (1e-9 rel / 1e-12 abs).
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--devices", type=int, default=0, help="GPUs of this process (0: all); ignored under torchrun")
    ap.add_argument("--nuclides", type=int, default=24)
    ap.add_argument("--tile-rows", type=int, default=1024)
    ap.add_argument("--plan", default="lpt", choices=["lpt", "static"])
    ap.add_argument("--check", type=int, default=1, help="nuclides checked against the CPU oracle on the root")
    ap.add_argument("--ne-hi", type=int, default=40000)
    args = ap.parse_args()

    from ndpp_b200 import library
    from ndpp_b200.group import Group

    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = dev = None
    if world > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dev = torch.device("cuda", local)
        dist.init_process_group("nccl", device_id=dev)
        group = Group.from_rank(local, dist.get_rank(), world, library.broadcast_id(dist, dev))
    else:
        group = Group(args.devices)
    keep = None
    if args.check > 0:      # the CPU oracle is the checker, never the thing measured (tests/util.py)
        from tests.util import check_library_against_oracle
        keep = check_library_against_oracle(args.check)
    out = library.run_c5(group, args.nuclides, dist=dist, policy=args.plan, tile_rows=args.tile_rows, ne_hi=args.ne_hi,
                         keep=keep, device=dev)
    if out is not None:
        print(json.dumps(out))
    group.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
