#!/bin/bash
# ncu --set full capture of one H.v kernel launch: tools/ncu_apply.sh <workload> <mode of tools/apply_fold_ab.py> <out name>
set -e
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:apply_" -s 1 -c 1 -f -o gpurun_out/$3 \
    python tools/apply_fold_ab.py $1 --reps 1 --modes $2 > gpurun_out/$3.log 2>&1
ncu -i gpurun_out/$3.ncu-rep --page raw --csv > gpurun_out/$3_raw.csv 2>/dev/null
ncu -i gpurun_out/$3.ncu-rep --page source --csv > gpurun_out/$3_source.csv 2>/dev/null || true
rm -f gpurun_out/$3.ncu-rep
