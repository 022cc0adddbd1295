#!/bin/bash
# A/B by environment: ws (default), solo, legacy
for v in ws solo legacy; do
  export NDPPGPU_F6_SOLO=0 NDPPGPU_F6_LEGACY=0
  [ $v = solo ] && export NDPPGPU_F6_SOLO=1
  [ $v = legacy ] && export NDPPGPU_F6_LEGACY=1
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<P
import json
try:
    d=json.load(open("gpurun_out/bench_ab_$v.json")); r=d["roofline"]
    print("$v value %.4g e2e %.4g ms/step %.1f f6_ms %.1f frac %.3f"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["kernel_ms"],r["frac"]))
except Exception as e: print("$v bench failed", e)
P
done
