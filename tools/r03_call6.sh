#!/bin/bash
# Session-3 GPU call 6: rows kernel, register-term variant (REGT) vs shared-memory extras; heavy path compiled out when unused.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "rows_kernel or lanes_kernel_clusters or top_of_32 or large_G_default" > gpurun_out/r03_rows6_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r03_rows6_pytest.log
tail -5 gpurun_out/r03_rows6_pytest.log
S=gpurun_out/r03_rows6_sweep.jsonl; : > $S
E=gpurun_out/r03_sweep6_err.log; : > $E
timeout 300 python tools/fill_sweep.py C3 --rows 18 --max-gb 10 --reps 10 --cfgs "auto rows:0:1:8 rows:0:1:9" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H8 --reps 10 --cfgs "auto rows:0 rows:1:1:5 rows:1:1:6 rows:1:1:7 rows:1:1:8 rows:1:1:7:6:5 rows:1:1:7:6:6" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H12 --rows 18 --reps 10 --cfgs "auto rows:0 rows:1:1 rows:1:2 rows:1:2:7 rows:1:1:7 rows:1:2:9" >> $S 2>>$E
timeout 200 python tools/fill_sweep.py H6 --reps 20 --cfgs "lanes rows:1:2 rows:0:2" >> $S 2>>$E
cat $S
tail -3 $E
