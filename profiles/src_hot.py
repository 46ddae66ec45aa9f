#!/usr/bin/env python
"""Per-source-line hot spots from an .ncu-rep captured with --import-source on
(kernels built with -lineinfo).  Usage: src_hot.py report.ncu-rep [kernel-regex] [top-n]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kernel = sys.argv[2] if len(sys.argv) > 2 and sys.argv[2] != "-" else None
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"]
if kernel:
    cmd += ["-k", f"regex:{kernel}"]
rows = list(csv.reader(io.StringIO(subprocess.run(cmd, capture_output=True, text=True).stdout)))
fpath, func, hdr, lines = None, None, None, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fpath = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        func = r[1]
    elif r[0] == "Line No":
        hdr = r
    elif hdr and r[0].strip().isdigit():
        ci, cs = hdr.index("Instructions Executed"), hdr.index("# Samples")
        try:
            key = (func, fpath, int(r[0]), r[1].strip()[:90])
            n, s = int(r[ci]), int(r[cs])
        except (ValueError, IndexError):
            continue
        a = lines.setdefault(key, [0, 0])
        a[0] += n; a[1] += s
by_func = {}
for (f, *_), (n, s) in lines.items():
    t = by_func.setdefault(f, [0, 0]); t[0] += n; t[1] += s
for f, (tn, ts) in by_func.items():
    print("#" * 100)
    print(f"{f[:95]}   warp-instr {tn / 1e6:.1f}M  samples {ts}")
    sel = [(v, k) for k, v in lines.items() if k[0] == f]
    print("  -- by samples")
    for (n, s), k in sorted(sel, key=lambda x: -x[0][1])[:top]:
        print(f"  {100 * s / max(ts, 1):5.1f}% smp {100 * n / max(tn, 1):5.1f}% ins  {k[1]}:{k[2]:<4d} {k[3]}")
    print("  -- by instructions")
    for (n, s), k in sorted(sel, key=lambda x: -x[0][0])[:top]:
        print(f"  {100 * n / max(tn, 1):5.1f}% ins {100 * s / max(ts, 1):5.1f}% smp  {k[1]}:{k[2]:<4d} {k[3]}")
