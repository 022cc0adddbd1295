#!/bin/bash
python -m pytest tests -m gpu -q 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
timeout 900 python scripts/run_configs.py --tag final2 > gpurun_out/configs_final2.log 2>&1; echo "configs rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_freegas_items -c 1 -f -o gpurun_out/prof_freegas_items python scripts/profile_case.py --case c3 --n 200 > gpurun_out/ncu_freegas_items.log 2>&1; echo "ncu rc=$?"
