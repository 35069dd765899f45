#!/bin/bash
# ncu --set full capture of one fill-kernel launch: tools/ncu_fill.sh <workload> <cfg> <kernel regex> <out name>
# e.g. tools/ncu_fill.sh C3 lanes:5:32 fill_lanes r02_lanes_C3 [log2 rows, default 17]
set -e
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k "regex:$3" -s 5 -c 1 -f -o gpurun_out/$4 \
    python tools/fill_sweep.py $1 --rows ${5:-17} --cfgs "$2" --reps 5 > gpurun_out/$4.log 2>&1
ncu -i gpurun_out/$4.ncu-rep --page raw --csv > gpurun_out/$4_raw.csv 2>/dev/null
