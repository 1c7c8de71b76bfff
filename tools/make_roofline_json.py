#!/usr/bin/env python3
"""profiles/roofline_traffic.json from `ncu --set full` captures of the SHIPPED kernels.

  python tools/make_roofline_json.py <ncu_raw.csv> [<ncu_raw.csv> ...]

Each CSV is the `ncu -i x.ncu-rep --page raw --csv` dump of one capture (any number of kernel rows).  For every
kernel row the tool records name, duration, DRAM bytes (read + write), tensor-pipe activity, issued instructions;
the file is stamped with the hash of the sources the library is built from (efficient_tts_b200.build.source_hash)
and the git revision.  bench.py prints these values next to its live CUDA-event timing only when the hash matches
the library it runs and the kernel name matches the instantiation the library reports for the tagged launch --
otherwise it prints nulls and says why.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficient_tts_b200 import build as B  # noqa: E402

FIELDS = {
    "duration_ms": "gpu__time_duration.sum",
    "dram_read": "dram__bytes_read.sum",
    "dram_write": "dram__bytes_write.sum",
    "tensor_pipe_active_pct": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "inst_executed": "smsp__inst_executed.sum",
    "sm_cycles_elapsed_max": "sm__cycles_elapsed.max",
    "dram_throughput_pct": "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
}
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}


def short_name(full):
    n = full.replace("void ", "").replace("efts::", "").replace("(int)", "")
    return n[:n.index("(")] if "(" in n else n


def parse(path):
    rows = list(csv.reader(open(path, newline="")))
    head, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(head)}
    out = []
    for r in rows[2:]:
        if len(r) < len(head):
            continue
        rec = {"kernel": short_name(r[col["Kernel Name"]]), "grid": r[col["Grid Size"]], "block": r[col["Block Size"]]}
        for key, metric in FIELDS.items():
            if metric not in col:
                continue
            raw = r[col[metric]].replace(",", "")
            if raw in ("", "no data", "n/a"):
                continue
            rec[key] = float(raw) * SCALE.get(units[col[metric]], 1.0)
        if "dram_read" in rec and "dram_write" in rec:
            rec["dram_bytes"] = rec["dram_read"] + rec["dram_write"]
        out.append(rec)
    return out


def main():
    if len(sys.argv) < 2:
        raise SystemExit(__doc__)
    kernels = []
    for p in sys.argv[1:]:
        for rec in parse(p):
            rec["source"] = os.path.relpath(os.path.abspath(p), ROOT)
            kernels.append(rec)
    try:
        rev = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], stdout=subprocess.PIPE,
                             text=True).stdout.strip()
    except Exception:
        rev = None
    res = {"source_sha16": B.source_hash(), "git_rev_at_generation": rev,
           "how": "ncu --set full --clock-control none, one launch per kernel; dram_bytes = dram__bytes_read.sum + "
                  "dram__bytes_write.sum; tensor_pipe_active_pct = sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "kernels": kernels}
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(res, f, indent=1)
    for k in kernels:
        print(k["kernel"], {x: k.get(x) for x in ("duration_ms", "dram_bytes", "tensor_pipe_active_pct")})


if __name__ == "__main__":
    main()
