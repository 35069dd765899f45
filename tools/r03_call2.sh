#!/bin/bash
# Session-3 GPU call 2: rows kernel with the CTA-wide heavy-group path: parity, sweep on the molecular operators.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "rows_kernel or lanes_kernel_clusters or top_of_32 or large_G_default" > gpurun_out/r03_rows2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03_rows2_pytest.log
tail -5 gpurun_out/r03_rows2_pytest.log
S=gpurun_out/r03_rows2_sweep.jsonl; : > $S
E=gpurun_out/r03_sweep2_err.log; : > $E
timeout 300 python tools/fill_sweep.py C3 --rows 18 --max-gb 10 --reps 10 --cfgs "rows:512:1:6 rows:512:1:7 rows:512:1:8 auto" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H8 --reps 10 --cfgs "lanes rows:512:1:7:0 rows:512:1:7:4 rows:512:1:7:6 rows:512:1:7:12 rows:1024:1:7:6 rows:512:1:5:6 auto" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "lanes rows:512:1:7:4 rows:512:1:7:6 rows:512:1:7:12 rows:1024:1:7:6 auto" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H6 --reps 20 --cfgs "lanes rows:512:2:7:6 rows:512:1:7:6 auto" >> $S 2>>$E
cat $S
tail -3 $E
