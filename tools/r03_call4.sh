#!/bin/bash
# Session-3 GPU call 4: rows kernel after the instruction diet (one POPC per term and batch, fma(+-1, c', acc)), adaptive run length.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "rows_kernel or lanes_kernel_clusters or top_of_32 or large_G_default" > gpurun_out/r03_rows4_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03_rows4_pytest.log
tail -5 gpurun_out/r03_rows4_pytest.log
S=gpurun_out/r03_rows4_sweep.jsonl; : > $S
E=gpurun_out/r03_sweep4_err.log; : > $E
timeout 300 python tools/fill_sweep.py C3 --rows 18 --max-gb 10 --reps 10 --cfgs "auto rows:512:1:8" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H8 --reps 10 --cfgs "auto rows:512 rows:1024:1:6" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "auto rows:512" >> $S 2>>$E
cat $S
tail -3 $E
QR_FILL_ROWS_HVS=5 timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "auto" | sed "s/auto/auto_hvs5/" >> $S
QR_FILL_ROWS_HVS=6 timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "auto" | sed "s/auto/auto_hvs6/" >> $S
tail -2 $S
