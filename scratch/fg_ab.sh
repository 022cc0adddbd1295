#!/bin/bash
for v in "$@"; do
  cp scratch/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  timeout 300 python scratch/fg_ab.py $v 2>&1 | tail -1
done
cp scratch/libs/fgs4.so ndpp_b200/csrc/libndppgpu.so
timeout 300 python -m pytest tests -m gpu -x -q -k "freegas or c3" 2>&1 | tail -2
