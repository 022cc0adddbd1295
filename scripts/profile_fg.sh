#!/bin/bash
# ncu evidence for the free-gas kernel (C3, 1000 E_in, 293.6 K): a plain run first, then the full-set capture of the
# first (root) generation with the source page, then the launch list.  Reports are exported to CSV on the box
# (gpurun brings back at most 64 MiB).  Usage: gpurun -- 'bash scripts/profile_fg.sh TAG'
TAG=${1:-fg}
O=gpurun_out
mkdir -p $O
python scripts/profile_case.py --case c3 || exit 1
ncu --clock-control none --set full --import-source on -k regex:k_freegas_items -c 1 -o $O/${TAG}_ncu python scripts/profile_case.py --case c3 > $O/${TAG}_ncu.log 2>&1
ncu -i $O/${TAG}_ncu.ncu-rep --page raw --csv > $O/${TAG}_ncu.raw.csv 2>/dev/null
ncu -i $O/${TAG}_ncu.ncu-rep --page source --csv > $O/${TAG}_ncu.source.csv 2>/dev/null
rm -f $O/${TAG}_ncu.ncu-rep
ncu --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum -c 120 --csv --log-file $O/${TAG}_launches.csv python scripts/profile_case.py --case c3 > $O/${TAG}_launches.log 2>&1
ls -la $O | grep $TAG
