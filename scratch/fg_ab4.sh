#!/bin/bash
for v in fgi6 fgi5 fgi4; do
  cp scratch/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  for sp in 2 1; do NDPPGPU_FG_QUEUE=24000000 NDPPGPU_FG_SPLIT=$sp timeout 200 python scratch/fg_ab.py ${v}_split$sp 2>&1 | tail -1; done
done
cp scratch/libs/fgi6.so ndpp_b200/csrc/libndppgpu.so
NDPPGPU_FG_SPLIT=2 timeout 300 python scratch/fg_scale.py 2>&1 | head -6
