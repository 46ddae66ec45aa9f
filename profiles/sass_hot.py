#!/usr/bin/env python
"""Opcode mix and hot address regions from `ncu --page source --print-source sass --csv`."""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kernel = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else None)
cmd = ["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"]
if kernel:
    cmd += ["-k", f"regex:{kernel}"]
out = subprocess.run(cmd, capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
start = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
for si, s in enumerate(start):
    hdr = rows[s]
    end = start[si + 1] - 1 if si + 1 < len(start) else len(rows)
    data = rows[s + 1:end]
    print("#" * 80)
    print(rows[s - 1][:2] if s > 0 else "")
    ci, cs, cn = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    tot = stot = 0
    byop, sop, ex = collections.Counter(), collections.Counter(), []
    for r in data:
        try:
            n, smp = int(r[ci]), int(r[cn])
        except (ValueError, IndexError):
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[cs])
        op = m.group(2).split(".")[0] if m else "?"
        byop[op] += n; sop[op] += smp; tot += n; stot += smp
        ex.append((n, smp, r[cs]))
    print(f"total warp-instr {tot / 1e6:.1f}M, samples {stot}")
    for op, n in byop.most_common(16):
        print(f"  {op:10s} {n / 1e6:9.1f}M {100 * n / max(tot, 1):5.1f}%   samples {100 * sop[op] / max(stot, 1):5.1f}%")
    # hot windows of 32 instructions
    win = 48
    sums = [(sum(e[1] for e in ex[i:i + win]), i) for i in range(0, len(ex), win)]
    for smp, i in sorted(sums, reverse=True)[:6]:
        n = sum(e[0] for e in ex[i:i + win])
        print(f"  window @{i:6d}: samples {100 * smp / max(stot, 1):5.1f}%  exec {100 * n / max(tot, 1):5.1f}%  e.g. {ex[i][2][:50]}")
