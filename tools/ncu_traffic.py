#!/usr/bin/env python
"""Writes profiles/fill_traffic.json from an `ncu --set full --page raw --csv` export of ONE fill-kernel launch
(tools/ncu_fill.sh): dram__bytes_read.sum + dram__bytes_write.sum per launch, stamped with the hash of the fill
source it was taken from -- bench.py quotes it as roofline.traffic only while the hash matches.
  python tools/ncu_traffic.py gpurun_out/r05_staged_C2_raw.csv xxz_periodic_n20_J1_delta0.7 [profiles/<copy of the csv>]"""
import csv, hashlib, json, shutil, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent

UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def fill_source_hash():
    h = hashlib.sha256()
    for f in ("fill.cuh", "plan.cuh", "scan.cuh"):
        h.update((ROOT / "qrusty_b200" / "csrc" / f).read_bytes())
    return h.hexdigest()[:16]


def main():
    raw, workload = Path(sys.argv[1]), sys.argv[2]
    rows = list(csv.reader(open(raw)))
    hdr, units, line = rows[0], rows[1], rows[2]

    def val(name):
        i = hdr.index(name)
        return float(line[i].replace(",", "")) * UNIT[units[i]]

    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    kernel = line[hdr.index("Kernel Name")]
    if kernel.startswith("void "): kernel = kernel[5:]
    dst = raw
    if len(sys.argv) > 3:
        dst = Path(sys.argv[3]); shutil.copyfile(raw, dst)
    out = {
        "workload": workload,
        "kernel": kernel,
        "dram_bytes_per_launch": int(rd + wr),
        "dram_bytes_read": int(rd),
        "dram_bytes_write": int(wr),
        "fill_source_sha256_16": fill_source_hash(),
        "source": "%s (ncu --set full --clock-control none, one launch of %s on the full build)" % (dst, kernel.split("(")[0]),
        "note": "below the algorithmic bytes when the tail of the output is still dirty in the 126 MB L2 at kernel end",
    }
    (ROOT / "profiles" / "fill_traffic.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
