#!/usr/bin/env python3
"""Per-source-line warp-instruction counts and stall samples of an .ncu-rep, in line order.
usage: tools/ncu_lines.py report.ncu-rep [min_fraction]"""
import csv, io, subprocess, sys
rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
src = list(csv.reader(io.StringIO(out)))
sections = [i for i, r in enumerate(src) if r and r[0] == "Line No"]
seen, files = set(), []
for si, start in enumerate(sections):
    fpath = src[start - 2][1] if start >= 2 else ""
    if fpath in seen or "abdpymc" not in fpath:
        continue
    seen.add(fpath)
    h = src[start]
    end = sections[si + 1] - 2 if si + 1 < len(sections) else len(src)
    iL, iS, iI = h.index("Line No"), h.index("# Samples"), h.index("Instructions Executed")
    agg, text = {}, {}
    for r in src[start + 1:end]:
        if len(r) <= iI:
            continue
        try:
            ln = int(r[iL])
        except ValueError:
            continue
        a = agg.setdefault(ln, [0, 0])
        a[0] += int(r[iS]) if r[iS].isdigit() else 0
        a[1] += int(r[iI]) if r[iI].isdigit() else 0
        text[ln] = r[1]
    files.append((fpath, agg, text))
total = sum(a[1] for _, agg, _ in files for a in agg.values())
print(f"total warp instructions {total / 1e6:.2f} M")
for fpath, agg, text in files:
    print("---", fpath, f"{sum(a[1] for a in agg.values()) / 1e6:.2f} M")
    for ln in sorted(agg):
        if agg[ln][1] > total * thr:
            print(f"{ln:5d} s{agg[ln][0]:5d} i{agg[ln][1] / 1e6:7.2f}M {100 * agg[ln][1] / total:5.1f}%  {text[ln].strip()[:100]}")
