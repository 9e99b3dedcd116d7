#!/usr/bin/env python3
"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by (file, line range) regions.
usage: ncu_regions.py dump.csv file:lo-hi[:label] ..."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
regions = []
for a in sys.argv[2:]:
    parts = a.split(":")
    f = parts[0]; lo, hi = map(int, parts[1].split("-")); label = parts[2] if len(parts) > 2 else a
    regions.append((f, lo, hi, label))
fname = ""; hdr = None; data = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if '# Samples' in r:
        hdr = r; iS = r.index('# Samples'); iI = r.index('Instructions Executed'); iT = r.index('Thread Instructions Executed') if 'Thread Instructions Executed' in r else None; continue
    if hdr is None or len(r) <= iI or not r[0].strip().isdigit():
        continue
    try:
        data.append((fname, int(r[0]), int(r[iS]), int(r[iI]), int(r[iT]) if iT is not None else 0))
    except ValueError:
        pass
ts = sum(d[2] for d in data) or 1; ti = sum(d[3] for d in data) or 1; tt = sum(d[4] for d in data) or 1
print(f"total samples {ts} warp-instr {ti} thread-instr {tt}")
acc = {r[3]: [0, 0, 0] for r in regions}; other = [0, 0, 0]
for f, l, s, i, t in data:
    for rf, lo, hi, label in regions:
        if f == rf and lo <= l <= hi:
            acc[label][0] += s; acc[label][1] += i; acc[label][2] += t; break
    else:
        other[0] += s; other[1] += i; other[2] += t
for label, (s, i, t) in acc.items():
    print(f"{label:28s} {100*s/ts:5.1f}% smp {100*i/ti:5.1f}% ins  avg-active {t/max(i,1):5.1f}")
print(f"{'other':28s} {100*other[0]/ts:5.1f}% smp {100*other[1]/ti:5.1f}% ins  avg-active {other[2]/max(other[1],1):5.1f}")
