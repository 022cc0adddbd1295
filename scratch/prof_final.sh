#!/bin/bash
python scripts/run_library.py --nuclides 300 --check 0 > gpurun_out/lib300_n1_v2.json 2> gpurun_out/lib300_n1_v2.err
tail -c 500 gpurun_out/lib300_n1_v2.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch_final.log 2>&1
echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_freegas_warp -c 1 -f -o gpurun_out/prof_freegas_final python scripts/profile_case.py --case c3 --n 200 > gpurun_out/ncu_freegas_final.log 2>&1
echo "freegas ncu rc=$?"
ls -la gpurun_out/prof_freegas_final.ncu-rep
