#!/bin/bash
# final pass of this session: GPU tests, smoke, both bench arms, ncu launch list and one full capture of the dominant kernel
s=$(date +%s)
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "pytest took $(( $(date +%s) - s )) s"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_v10.json 2> gpurun_out/bench_v10.err; echo "bench rc=$?"
python bench.py --impl reference > gpurun_out/bench_v10_ref.json 2> gpurun_out/bench_v10_ref.err; echo "ref rc=$?"
cut -c1-330 gpurun_out/bench_v10.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_v10.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_v10.log 2>&1
echo "launch list rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_file6_cm_ws -c 1 -f -o gpurun_out/prof_f6ws_v10 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full_v10.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/prof_f6ws_v10.ncu-rep
