#!/bin/bash
# A/B of the decoupled rows kernel (QR_FILL_ROWS_DEC=1) against the batch-barrier one on one box.
mkdir -p gpurun_out
S=gpurun_out/${TAG:-r06}_dec_ab.jsonl; : > $S
for rep in 1; do
for dec in ${DECS:-0 1}; do
  export QR_FILL_ROWS_DEC=$dec
  for w in "H8" "H12 --rows 18" "H10 --rows 17" "H11 --rows 16" "C3 --rows 18 --max-gb 10" "rand:22:96:64 --rows 20" "rand:24:6000:3000 --rows 16" "xxz27 --rows 23"; do
    timeout 200 python tools/fill_sweep.py $w --reps 20 --cfgs "auto" 2>/dev/null | sed "s|\"cfg\": \"auto\"|\"cfg\": \"auto dec=$dec\"|" >> $S
  done
  for x in 0 1; do
    QR_FILL_ROWS_EXTHV=$x timeout 200 python tools/fill_sweep.py H10 --rows 17 --reps 20 --cfgs "rows::2:::::0:1024" 2>/dev/null | sed "s|\"cfg\": \"|\"cfg\": \"dec=$dec exthv=$x |" >> $S
  done
done
done
cut -c1-75,100-260 $S
