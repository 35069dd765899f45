#!/bin/bash
# Session-3 GPU call 3: rows kernel after the instruction diet (one POPC per term and batch, fma(+-1, c', acc)), adaptive run length.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "rows_kernel or lanes_kernel_clusters or top_of_32 or large_G_default" > gpurun_out/r03_rows3_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03_rows3_pytest.log
tail -5 gpurun_out/r03_rows3_pytest.log
S=gpurun_out/r03_rows3_sweep.jsonl; : > $S
E=gpurun_out/r03_sweep3_err.log; : > $E
timeout 300 python tools/fill_sweep.py C3 --rows 18 --max-gb 10 --reps 10 --cfgs "auto rows:512:1:8 rows:1024:1:8 rows:512:1:9" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H8 --reps 10 --cfgs "lanes auto rows:512 rows:1024 rows:1024:1:5 rows:1024:1:6 rows:1024:1:7" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "lanes auto rows:512 rows:1024 rows:1024:1:7 rows:1024:1:8:4" >> $S 2>>$E
cat $S
tail -3 $E
