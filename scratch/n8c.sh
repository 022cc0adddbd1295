#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29711 scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300v4_n8.json 2> gpurun_out/lib300v4_n8.err; echo "n8 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29712 scripts/run_library.py --nuclides 300 --check 0 --phases > gpurun_out/lib300v4_n8_phases.json 2> gpurun_out/lib300v4_n8_phases.err; echo "n8 phases rc=$?"
for f in n8 n8_phases; do tail -c 900 gpurun_out/lib300v4_$f.json; echo; done
