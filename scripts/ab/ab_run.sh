#!/bin/bash
# A/B of library variants built by scripts/ab/ab_build.py on the bench workload (C2): copies every variant over
# ndpp_b200/csrc/libndppgpu.so in turn, runs bench.py without the extras and prints the dominant kernel's ms; then runs
# the bit-identity / parity subset of the GPU tests on $CHECK.  Usage (on the GPU box):
#   VARIANTS="base shfl base" CHECK=shfl bash scripts/ab/ab_run.sh
VARIANTS=${VARIANTS:-"base"}
mkdir -p gpurun_out
cp ndpp_b200/csrc/libndppgpu.so /tmp/libndppgpu.default.so
for v in $VARIANTS; do
  cp scripts/ab/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  timeout 120 python bench.py --steps ${STEPS:-3} --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_ab_$v.json 2> gpurun_out/bench_ab_$v.err
  python - <<P
import json
try:
    d=[json.loads(l) for l in open("gpurun_out/bench_ab_$v.json") if l.startswith("{")][0]; r=d["roofline"]
    print("$v value %.4g e2e %.4g ms/step %.2f f6_ms %.2f frac %.3f clocks %s"%(d["value"],d["e2e"]["value"],d["ms_per_step"],r["kernel_ms"],r["frac"],d["clocks"]["sm_mhz"]))
except Exception as e: print("$v bench failed", e)
P
done
if [ -n "$CHECK" ]; then
  cp scripts/ab/libs/$CHECK.so ndpp_b200/csrc/libndppgpu.so
  timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q -k "bit_identical or c2_full or heavy_shape or c1_moments or law61 or legendre_leaf or golden or law9 or determin" 2>&1 | tail -3
fi
cp /tmp/libndppgpu.default.so ndpp_b200/csrc/libndppgpu.so
