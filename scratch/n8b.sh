#!/bin/bash
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port 29611 scripts/run_library.py --nuclides 300 --check 3 > gpurun_out/lib300v3_n8.json 2> gpurun_out/lib300v3_n8.err; echo "n8 rc=$?"
timeout 300 $TR --nproc-per-node 4 --master-port 29612 scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300v3_n4.json 2> gpurun_out/lib300v3_n4.err; echo "n4 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29613 scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300v3_n2.json 2> gpurun_out/lib300v3_n2.err; echo "n2 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29614 scripts/run_library.py --nuclides 300 --check 0 --plan static > gpurun_out/lib300v3_n8_static.json 2> gpurun_out/lib300v3_n8_static.err; echo "n8 static rc=$?"
for f in n8 n4 n2 n8_static; do tail -c 330 gpurun_out/lib300v3_$f.json; echo; done
