#!/bin/bash
# bench step (canonicalise + fill) with and without programmatic dependent launch
for pdl in 1 0; do
  QR_PDL=$pdl timeout 300 python bench.py --steps 20 --warmup 3 --no-hv --no-e2e --no-extras --no-cpu-baseline > gpurun_out/pdl_$pdl.json 2>gpurun_out/pdl_$pdl.err
  python - $pdl <<'PY'
import sys, json
pdl = sys.argv[1]
d = json.loads(open("gpurun_out/pdl_%s.json" % pdl).read().strip().splitlines()[-1])
print("QR_PDL=%s ms_per_step %.5f fill_ms %.5f launches %d runs %s" % (pdl, d["ms_per_step"], d["roofline"]["kernel_ms"], d["gpu_launches"], d["run"]["ms_per_step_runs"]))
PY
done
