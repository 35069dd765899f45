#!/bin/bash
# One GPU call that re-verifies a session: full suite + smoke + bench (own arm, reference arm) + launch list + ncu --set full of the
# headline fill kernel on C2 (-> profiles/fill_traffic.json via tools/ncu_traffic.py) and of the rows kernel on H12 / H8.
T=${TAG:-rNN}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -x -m gpu > gpurun_out/${T}_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/${T}_pytest_gpu.log
tail -4 gpurun_out/${T}_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; tail -2 gpurun_out/${T}_smoke.log
timeout 600 python bench.py > gpurun_out/${T}_bench_n1.json 2> gpurun_out/${T}_bench_n1.err; tail -c 3000 gpurun_out/${T}_bench_n1.json; tail -3 gpurun_out/${T}_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_reference_n1.json 2>> gpurun_out/${T}_bench_n1.err; tail -c 600 gpurun_out/${T}_bench_reference_n1.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${T}_launches_bench.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/${T}_launches_bench.log 2>&1
timeout 400 bash tools/ncu_fill.sh C2 auto fill_staged ${T}_staged_C2 20
python tools/ncu_traffic.py gpurun_out/${T}_staged_C2_raw.csv xxz_periodic_n20_J1_delta0.7 && cp profiles/fill_traffic.json gpurun_out/${T}_fill_traffic.json
if [ -z "$QUICK" ]; then
timeout 400 bash tools/ncu_fill.sh H12 auto fill_rows ${T}_rows_H12 16
timeout 400 bash tools/ncu_fill.sh H8 auto fill_rows ${T}_rows_H8 16
fi
rm -f gpurun_out/*.ncu-rep
ls -la gpurun_out | tail -12
