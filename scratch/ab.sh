#!/bin/bash
for v in "$@"; do
  cp scratch/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  timeout 90 python -m pytest tests -m gpu -x -q -k "ws_bit_identical" 2>&1 | tail -1
  timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$v value %.4g e2e %.4g f6_ms %.1f frac %.3f'%(d['value'],d['e2e']['value'],r['kernel_ms'],r['frac']))"
done
