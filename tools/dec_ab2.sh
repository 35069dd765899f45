#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/${TAG:-r06}_dec_ab.jsonl; : > $S
for dec in 1; do
  export QR_FILL_ROWS_DEC=$dec
  for w in "H10 --rows 17" "H11 --rows 16" "rand:24:6000:3000 --rows 16"; do
    for perm in 0 1 ""; do
    QR_FILL_ROWS_PERM=$perm timeout 200 python tools/fill_sweep.py $w --reps 20 --cfgs "auto" 2>/dev/null | sed "s|\"cfg\": \"auto\"|\"cfg\": \"auto dec=$dec perm=$perm\"|" >> $S
    done
  done
done
cut -c1-75,100-260 $S
