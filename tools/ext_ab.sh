#!/bin/bash
mkdir -p gpurun_out
S=gpurun_out/${TAG:-r06}_ext_ab.jsonl; : > $S
for w in "H8" "H12 --rows 18" "H6" "rand:20:3000:600 --rows 17"; do
  for x in 0 1 ""; do for dec in 0 1; do
    QR_FILL_ROWS_DEC=$dec QR_FILL_ROWS_EXTHV=$x timeout 200 python tools/fill_sweep.py $w --reps 20 --cfgs "auto" 2>/dev/null | sed "s|\"cfg\": \"auto\"|\"cfg\": \"auto exthv=$x dec=$dec\"|" >> $S
  done; done
done
cut -c1-75,100-260 $S
