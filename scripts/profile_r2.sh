#!/bin/bash
# Round-2 ncu evidence (run on the GPU box through gpurun; the summaries under profiles/ are made from the reports
# with scripts/summarize_ncu.py).  Every ncu pass follows an untimed plain run of the same command that exited 0.
set -u
O=gpurun_out
mkdir -p $O
NCU="ncu --clock-control none"
python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_prof_plain.json 2> $O/r2_prof_plain.err || exit 1
# (a) launch list of the bench command
$NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $O/r2_launches_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_launches_c2.log 2>&1
# (b) the dominant kernel at full size (the launch of the timed step: skip the 3 warm-up launches)
$NCU --set full --import-source on -k regex:k_file6_cm_ws -s 3 -c 1 -o $O/r2_ncu_file6_cm_ws_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r2_ncu_f6.log 2>&1
# (c) the other kernels of the path
python scripts/profile_case.py --case c2 --n 20000 || exit 1
$NCU --set full -k regex:'k_inelastic|k_elastic|k_unitbase|k_f6_femu|k_f6_records|k_file6_reduce|k_convert_file6' -c 12 -o $O/r2_ncu_c2_others python scripts/profile_case.py --case c2 > $O/r2_ncu_c2o.log 2>&1
python scripts/profile_case.py --case c3 || exit 1
$NCU --set full --import-source on -k regex:k_freegas_items -c 1 -o $O/r2_ncu_freegas_items_c3 python scripts/profile_case.py --case c3 > $O/r2_ncu_fg.log 2>&1
$NCU --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active -c 60 --csv --log-file $O/r2_launches_c3.csv python scripts/profile_case.py --case c3 > $O/r2_launches_c3.log 2>&1
python scripts/profile_case.py --case c4d && python scripts/profile_case.py --case c4c || exit 1
$NCU --set full -k regex:k_sab -c 6 -o $O/r2_ncu_sab_discrete python scripts/profile_case.py --case c4d > $O/r2_ncu_c4d.log 2>&1
$NCU --set full -k regex:k_sab -c 6 -o $O/r2_ncu_sab_continuous python scripts/profile_case.py --case c4c > $O/r2_ncu_c4c.log 2>&1
# the reports are large (gpurun brings back at most 64 MiB): keep their raw / source pages as CSV, drop the reports
for r in $O/r2_ncu_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
done
for r in $O/r2_ncu_file6_cm_ws_c2 $O/r2_ncu_freegas_items_c3; do
  ncu -i $r.ncu-rep --page source --csv > $r.source.csv 2>/dev/null
done
cuobjdump -sass -fun '_ZN4ndpp13k_file6_cm_wsILi8EEEvNS_6NucDevEPKdNS_5UbDevEPKNS_5UbRecEPKiS3_S3_S9_iiPyPd' ndpp_b200/csrc/libndppgpu.so > $O/r2_sass_k_file6_cm_ws_8.txt 2>/dev/null
rm -f $O/r2_ncu_*.ncu-rep
ls -la $O | tail -30
du -sh $O
