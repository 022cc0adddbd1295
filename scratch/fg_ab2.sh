#!/bin/bash
cp scratch/libs/fgs5.so ndpp_b200/csrc/libndppgpu.so; timeout 200 python scratch/fg_ab.py fgs5 2>&1 | tail -1
cp scratch/libs/fgi4.so ndpp_b200/csrc/libndppgpu.so
for sp in 4 3 2; do NDPPGPU_FG_SPLIT=$sp timeout 200 python scratch/fg_ab.py fgi4_split$sp 2>&1 | tail -1; done
timeout 300 python scratch/fg_scale.py 2>&1 | tail -12
timeout 600 python -m pytest tests -m gpu -x -q -k "freegas or c3 or golden or smoke" 2>&1 | tail -3
