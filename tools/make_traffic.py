"""profiles/traffic.json from the round's `ncu --set full` captures (run in the build container).

    python tools/make_traffic.py <basefc report.ncu-rep> <reads of the profiled k_basefc_count launch> [<baf report> <reads>]

DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum) of ONE launch of k_basefc_count (and k_baf_scan), the reads
that launch processed, the commit the capture was taken at.  bench.py scales bytes / read to its own launch size and
echoes commit / report / when next to `roofline.traffic`, so a stale file shows.
"""
import csv
import datetime
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def dram_bytes(rep, kernel):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kernel],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    r = rows[2]
    tot = 0.0
    for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
        tot += float(r[ci[m]].replace(",", "")) * UNIT[units[ci[m]]]
    tu = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[units[ci["gpu__time_duration.sum"]]]
    return tot, float(r[ci["gpu__time_duration.sum"]].replace(",", "")) * tu


out = {"commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], cwd=ROOT, capture_output=True, text=True).stdout.strip(),
       "when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%MZ"),
       "report": os.path.basename(sys.argv[1])}
b, ms = dram_bytes(sys.argv[1], "k_basefc_count")
out["k_basefc_count"] = {"dram_bytes": b, "reads": int(float(sys.argv[2])), "duration_under_ncu_ms": ms}
if len(sys.argv) > 4:
    b, ms = dram_bytes(sys.argv[3], "k_baf_scan")
    out["k_baf_scan"] = {"dram_bytes": b, "reads": int(float(sys.argv[4])), "duration_under_ncu_ms": ms, "report": os.path.basename(sys.argv[3])}
with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as fp:
    json.dump(out, fp, indent=1)
    fp.write("\n")
print(json.dumps(out, indent=1))
