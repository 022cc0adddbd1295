#!/bin/bash
for v in fghead fgpair fghead fgpair; do
  cp scratch/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  timeout 200 python scratch/fg_ab.py $v 2>&1 | tail -1
done
cp scratch/libs/fgpair.so ndpp_b200/csrc/libndppgpu.so
timeout 600 python -m pytest tests -m gpu -x -q -k "freegas or c3 or golden" 2>&1 | tail -3
