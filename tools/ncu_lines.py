#!/usr/bin/env python3
"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
fname = ""
data = []
hdr = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if '# Samples' in r:
        hdr = r; iS = r.index('# Samples'); iI = r.index('Instructions Executed'); continue
    if hdr is None or len(r) <= iI or not r[0].strip().isdigit():
        continue
    try:
        data.append((int(r[iS]), int(r[iI]), fname, int(r[0]), r[1].strip()))
    except ValueError:
        pass
ts = sum(d[0] for d in data) or 1; ti = sum(d[1] for d in data) or 1
print(f"total samples {ts}  total warp-instructions {ti}")
print("--- by stall samples")
for s, n, f, l, src in sorted(data, reverse=True)[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*n/ti:5.1f}% ins  {f}:{l:<4d} {src[:100]}")
print("--- by instructions")
for s, n, f, l, src in sorted(data, key=lambda d: -d[1])[:top]:
    print(f"{100*s/ts:5.1f}% smp {100*n/ti:5.1f}% ins  {f}:{l:<4d} {src[:100]}")
