#!/bin/bash
cp scratch/libs/fgi4.so ndpp_b200/csrc/libndppgpu.so
for sp in 2 1; do NDPPGPU_FG_QUEUE=40000000 NDPPGPU_FG_SPLIT=$sp timeout 200 python scratch/fg_ab.py fgi4_split$sp 2>&1 | tail -1; done
NDPPGPU_FG_QUEUE=40000000 NDPPGPU_FG_SPLIT=1 timeout 300 python scratch/fg_scale.py 2>&1 | tail -12
