#!/usr/bin/env python3
"""Stall reasons per CUDA source line (and in total) from an `ncu --page source --print-source cuda,sass --csv` dump (optionally .gz).
usage: ncu_stalls.py <csv[.gz]> [top]"""
import csv, gzip, sys
path = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
f = gzip.open(path, "rt") if path.endswith(".gz") else open(path)
hdr = None; fname = ""; tot = {}; lines = []
for r in csv.reader(f):
    if len(r) >= 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]; continue
    if "# Samples" in r:
        hdr = r; cols = [(i, c) for i, c in enumerate(r) if c.startswith("stall_") and "Not Issued" not in c]; iS = r.index("# Samples"); continue
    if hdr is None or not r or not r[0].strip().isdigit():
        continue
    try:
        st = {c: int(r[i]) for i, c in cols}; n = int(r[iS])
    except (ValueError, IndexError):
        continue
    for c, v in st.items():
        tot[c] = tot.get(c, 0) + v
    lines.append((n, fname, int(r[0]), st, r[1].strip()))
T = sum(tot.values()) or 1
print("total samples by reason:", {k[6:]: f"{100 * v / T:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v * 200 > T})
for n, fn, ln, st, src in sorted(lines, key=lambda x: -x[0])[:top]:
    main = sorted(st.items(), key=lambda kv: -kv[1])[:3]
    print(f"{100 * n / T:5.1f}%  {fn}:{ln:<5d} " + " ".join(f"{k[6:]}={100 * v / max(n, 1):.0f}%" for k, v in main) + "  | " + src[:70])
