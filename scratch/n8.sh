#!/bin/bash
# full C5 library at 8/4/2 GPUs, bench.py at 8 GPUs (round-1 scaling evidence)
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 400 $TR --nproc-per-node 8 --master-port 29511 scripts/run_library.py --nuclides 300 --check 3 > gpurun_out/lib300_n8.json 2> gpurun_out/lib300_n8.err
echo "lib n8 rc=$?"; tail -c 400 gpurun_out/lib300_n8.json
timeout 300 $TR --nproc-per-node 8 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_n8.json 2> gpurun_out/bench_n8.err
echo "bench n8 rc=$?"; head -c 300 gpurun_out/bench_n8.json
timeout 300 $TR --nproc-per-node 4 --master-port 29513 scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300_n4.json 2> gpurun_out/lib300_n4.err
echo "lib n4 rc=$?"
timeout 300 $TR --nproc-per-node 2 --master-port 29514 scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300_n2.json 2> gpurun_out/lib300_n2.err
echo "lib n2 rc=$?"
timeout 300 $TR --nproc-per-node 8 --master-port 29515 scripts/run_library.py --nuclides 300 --check 0 --plan static > gpurun_out/lib300_n8_static.json 2> gpurun_out/lib300_n8_static.err
echo "lib n8 static rc=$?"
