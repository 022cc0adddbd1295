#!/bin/bash
# usage: VARIANTS="a b" bash ab_ncu.sh  -- DRAM traffic of k_freegas_items per variant
cp ndpp_b200/csrc/libndppgpu.so /tmp/libndppgpu.default.so
for v in $VARIANTS; do
  cp scripts/ab/libs/$v.so ndpp_b200/csrc/libndppgpu.so
  ncu --clock-control none -k regex:k_freegas_items -c 1 --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,l1tex__t_sector_hit_rate.pct,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,smsp__inst_executed.sum --csv --log-file gpurun_out/abncu_$v.csv python scripts/profile_case.py --case c3 > /dev/null 2>&1
  echo $v; python - <<P
import csv
rows=[r for r in csv.reader(l for l in open("gpurun_out/abncu_$v.csv") if l.startswith('"'))]
h=rows[0]
for r in rows[1:]: print("   ", r[h.index("Metric Name")], r[h.index("Metric Value")], r[h.index("Metric Unit")])
P
done
cp /tmp/libndppgpu.default.so ndpp_b200/csrc/libndppgpu.so
