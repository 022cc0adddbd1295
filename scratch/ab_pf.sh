#!/bin/bash
# A/B of library variants in scratch/libs on the bench workload (C2); prints the dominant kernel's ms
VARIANTS=${VARIANTS:-"base pf pf_s4 base pf"}
for v in $VARIANTS; do
  cp scratch/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  timeout 90 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/bench_ab_$v.json")); r=d["roofline"]
    print("$v value %.4g e2e %.4g ms/step %.1f f6_ms %.1f frac %.3f clocks %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["kernel_ms"],r["frac"],d["clocks"]["sm_mhz"]))
except Exception as e: print("$v bench failed", e)
P
done
LAST=${CHECK:-pf}
cp scratch/libs/$LAST.so ndpp_b200/csrc/libndppgpu.so
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "bit_identical or c2_full or heavy_shape or c1_moments or law61 or legendre_leaf or golden or law9" 2>&1 | tail -3
