#!/bin/bash
# One GPU call that re-verifies a session: full suite + smoke + bench (own arm, reference arm) + launch list + ncu of the rows kernel on C3 and H12.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -x -m gpu > gpurun_out/${TAG:-rNN}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${TAG:-rNN}_pytest_gpu.log
tail -4 gpurun_out/${TAG:-rNN}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG:-rNN}_smoke.log 2>&1; tail -2 gpurun_out/${TAG:-rNN}_smoke.log
timeout 600 python bench.py > gpurun_out/${TAG:-rNN}_bench_n1.json 2> gpurun_out/${TAG:-rNN}_bench_n1.err; tail -c 3000 gpurun_out/${TAG:-rNN}_bench_n1.json; tail -3 gpurun_out/${TAG:-rNN}_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/${TAG:-rNN}_bench_reference_n1.json 2>> gpurun_out/${TAG:-rNN}_bench_n1.err; tail -c 600 gpurun_out/${TAG:-rNN}_bench_reference_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG:-rNN}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/${TAG:-rNN}_launches_bench.log 2>&1
timeout 400 bash tools/ncu_fill.sh C3 auto fill_rows ${TAG:-rNN}_rows_C3 16
timeout 400 bash tools/ncu_fill.sh H12 auto fill_rows ${TAG:-rNN}_rows_H12 16
timeout 400 bash tools/ncu_fill.sh H8 auto fill_rows ${TAG:-rNN}_rows_H8 16
ls -la gpurun_out | tail -20
