#!/bin/bash
# Session-3 GPU call 1: parity of the new rows kernel, sweep against the lanes kernel, one ncu capture.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "rows_kernel or lanes_kernel_clusters or top_of_32 or large_G_default" > gpurun_out/r03_rows_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03_rows_pytest.log
tail -5 gpurun_out/r03_rows_pytest.log
S=gpurun_out/r03_rows_sweep.jsonl; : > $S
timeout 300 python tools/fill_sweep.py C3 --rows 18 --max-gb 10 --reps 10 --cfgs "lanes rows:1024:1:5 rows:1024:1:7 rows:1024:1:9 rows:512:1:7 auto" >> $S 2>gpurun_out/r03_sweep_err.log
timeout 200 python tools/fill_sweep.py H8 --reps 10 --cfgs "lanes rows:1024:1:7 rows:512:1:7 auto" >> $S 2>>gpurun_out/r03_sweep_err.log
timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "lanes rows:1024:1:7 rows:512:1:7 auto" >> $S 2>>gpurun_out/r03_sweep_err.log
timeout 200 python tools/fill_sweep.py H10 --rows 17 --reps 5 --cfgs "lanes auto" >> $S 2>>gpurun_out/r03_sweep_err.log
cat $S
tail -3 gpurun_out/r03_sweep_err.log
timeout 400 bash tools/ncu_fill.sh C3 rows:1024:1:7 fill_rows r03_rows_C3 16
grep -i "gpu__time_duration.sum\|dram__bytes_write.sum\|dram__bytes_read.sum" gpurun_out/r03_rows_C3_raw.csv | head -3
ls -la gpurun_out | head -30
