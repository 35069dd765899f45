#!/bin/bash
# Probe of the staged fill's shared-memory layout for G a multiple of 4: no gaps (QR_FILL_PAD=0) against a
# 16-byte gap every 2 / 4 / 8 rows (QR_FILL_PAD=1/2/3), with other G as controls.
for w in xxz23 xxz15 xxz19 xxz27 xxz22 C2 C4; do
  for pad in 0 1 2 3; do QR_FILL_PAD=$pad python tools/fill_sweep.py $w --rows 20 --cfgs "2,8" | sed "s/^/pad=$pad /"; done
  python tools/fill_sweep.py $w --rows 20 --cfgs "auto" | sed "s/^/default /"
done
