#!/usr/bin/env python3
"""Opcode histogram of the shipped library, per kernel: `cuobjdump -sass libefts_b200.so` split by function.

  python tools/sass_histogram.py [out_prefix]      # default profiles/r02/sass_histogram -> .json and .md

The mnemonics that prove the Blackwell paths (B200_PROFILING.md): UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM
(tcgen05.ld / st), UTMALDG / UTMASTG (TMA bulk tensor load / store), UTCBAR (tcgen05.commit), SYNCS (mbarrier),
plus the plain tensor / memory opcodes a recompiled baseline would show instead (HMMA, LDG, STG, LDS, STS).
Runs on the build machine (no GPU needed); stamps the hash of the sources the library was built from.
"""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from efficient_tts_b200 import build as B  # noqa: E402

WATCH = ["UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "UTCCP",
         "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS", "LDSM", "ATOM", "RED", "MUFU", "FFMA", "FADD", "FMUL",
         "DADD", "DFMA", "SHFL", "BAR", "UCGABAR", "ACQBULK", "CCTL", "ERRBAR", "MEMBAR"]


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02", "sass_histogram")
    lib = B.build_library()
    sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["cu++filt"] + names, stdout=subprocess.PIPE, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    kernels = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            cur = re.sub(r"\(.*", "", cur.replace("(int)", "")).replace("void ", "").replace("efts::", "")
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
        if m and cur is not None:
            kernels[cur][m.group(1)] += 1
            kernels[cur]["_total"] += 1
    res = {"library": os.path.relpath(lib, ROOT), "source_sha16": B.source_hash(),
           "nvcc_flags": " ".join(B.NVCC_FLAGS), "kernels": {}}
    for k, c in kernels.items():
        res["kernels"][k] = {"instructions": c["_total"], **{w: c[w] for w in WATCH if c[w]}}
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    with open(prefix + ".json", "w") as f:
        json.dump(res, f, indent=1)
    cols = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "SYNCS", "HMMA", "LDG", "STG", "LDS", "STS"]
    with open(prefix + ".md", "w") as f:
        f.write("# SASS opcode histogram of `%s` (sources %s)\n\n" % (res["library"], res["source_sha16"]))
        f.write("`cuobjdump -sass`, instruction counts per kernel (static). UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, "
                "UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops.\n\n")
        f.write("| kernel | instr | " + " | ".join(cols) + " |\n|---|---|" + "---|" * len(cols) + "\n")
        for k, v in res["kernels"].items():
            f.write("| `%s` | %d | %s |\n" % (k, v["instructions"], " | ".join(str(v.get(c, 0)) for c in cols)))
        tot = collections.Counter()
        for v in res["kernels"].values():
            for c in cols:
                tot[c] += v.get(c, 0)
        f.write("| **total** | | %s |\n" % " | ".join(str(tot[c]) for c in cols))
    print(prefix + ".json", {c: sum(v.get(c, 0) for v in res["kernels"].values()) for c in cols[:6]})


if __name__ == "__main__":
    main()
