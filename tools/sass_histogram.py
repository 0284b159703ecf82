#!/usr/bin/env python3
"""SASS opcode histogram per kernel of libabd_b200.so (static instruction counts from cuobjdump -sass).
usage: python tools/sass_histogram.py > profiles/<round>_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
so = sys.argv[1] if len(sys.argv) > 1 else str(ROOT / "abdpymc_b200" / "libabd_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
hist, cur = collections.OrderedDict(), None
for ln in out.splitlines():
    m = re.search(r"Function : (\S+)", ln)
    if m:
        cur = m.group(1)
        hist[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
    if m and cur:
        hist[cur][m.group(1).split(".")[0]] += 1


def short(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = r.replace("(anonymous namespace)::", "")
    m = re.search(r"(k_\w+(?:<[^(]*>)?)", r)
    return m.group(1) if m else r[:80]


print("SASS opcode histogram per kernel of libabd_b200.so (cuobjdump -sass of the sm_100a cubin; STATIC instruction counts).")
print("UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier, DFMA / DMUL / DADD = fp64 pipe, MUFU = reciprocal seed,")
print("SHFL / REDUX = warp exchange, ATOMG / REDG = global atomics / reductions, LDL / STL = local-memory (spill) traffic.\n")
for k, c in hist.items():
    print(f"{short(k)}  [{sum(c.values())} instructions]")
    print("   " + "  ".join(f"{op} {n}" for op, n in c.most_common(28)))
    special = {op: c[op] for op in ("UBLKCP", "SYNCS", "MUFU", "ATOMG", "REDG", "LDL", "STL", "UTMALDG", "ACQBULK") if c[op]}
    print(f"   of note: {special}\n")
