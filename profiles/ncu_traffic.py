#!/usr/bin/env python
"""dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel, from `ncu --set full` reports -> the JSON bench.py
reads for `roofline.traffic` (profiles/r02_ncu_traffic.json), keyed by a hash of the kernel sources (bench.source_hash):
a figure captured on other sources is reported as null by the bench, not silently reused.

    python profiles/ncu_traffic.py gpurun_out/r2_prof_icp.ncu-rep gpurun_out/r2_prof_occ.ncu-rep ... [--command "..."]

Groups: icp_pairs_kernel = bulk + hand-over launches of one step (brute mode); icp_pairs_kernel_grid = the grid-mode
launch; occupancy_update = every kernel of one update."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT]
import bench  # noqa: E402

reps = [a for a in sys.argv[1:] if a.endswith(".ncu-rep")]
command = sys.argv[sys.argv.index("--command") + 1] if "--command" in sys.argv else ""
per_kernel = {}
for rep in reps:
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        units_r, units_w = rows[1][col["dram__bytes_read.sum"]], rows[1][col["dram__bytes_write.sum"]]
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        rd = float(r[col["dram__bytes_read.sum"]].replace(",", "")) * scale.get(units_r, 1.0)
        wr = float(r[col["dram__bytes_write.sum"]].replace(",", "")) * scale.get(units_w, 1.0)
        lts = float(r[col["lts__t_bytes.sum"]].replace(",", "")) * scale.get(rows[1][col["lts__t_bytes.sum"]], 1.0) if "lts__t_bytes.sum" in col else 0.0
        dur = float(r[col["gpu__time_duration.sum"]].replace(",", "")) if "gpu__time_duration.sum" in col else 0.0
        per_kernel.setdefault(name, []).append(dict(dram=rd + wr, lts=lts, ns=dur, grid=r[col["Grid Size"]] if "Grid Size" in col else ""))


def short(name):
    return name.split("(")[0].replace("icpb::", "").replace("<unnamed>::", "")


groups = {}
detail = {}
for name, launches in per_kernel.items():
    s = short(name)
    mean = sum(l["dram"] for l in launches) / len(launches)
    detail[s] = dict(launches=len(launches), dram_bytes_mean=mean, lts_bytes_mean=sum(l["lts"] for l in launches) / len(launches),
                     per_launch=[dict(dram=l["dram"], lts=l["lts"], ns=l["ns"]) for l in launches])
# one ICP step = one bulk launch + one hand-over launch: take the first captured launch of each of the brute kernels
import re
def variant(name):
    """(dim, grid, min blocks, threads) of an icp_pairs_kernel instantiation, however ncu spells the template arguments."""
    m = re.search(r"icp_pairs_kernel<\s*(?:\(int\))?(\d),\s*(?:\(bool\))?(\d|true|false),\s*(?:\(int\))?(\d),\s*(?:\(int\))?(\d+)(?:,\s*(?:\(int\))?\d+)?>", name)
    if not m:
        return None
    g = m.group(2)
    return (int(m.group(1)), 1 if g in ("1", "true") else 0, int(m.group(3)), int(m.group(4)))
bulk = [l for n, ls in per_kernel.items() if variant(n) and variant(n)[:2] == (2, 0) and variant(n)[2] > 1 for l in ls]
# (the hand-over goes out as two launches side by side, clusters and plain CTAs: the first launch of each variant)
hand = [ls[0] for n, ls in per_kernel.items() if variant(n) and variant(n)[:2] == (2, 0) and variant(n)[2] == 1]
grid = [(n, l) for n, ls in per_kernel.items() if variant(n) and variant(n)[1] == 1 for l in ls]
if bulk and hand:
    groups["icp_pairs_kernel"] = bulk[0]["dram"] + sum(l["dram"] for l in hand)
if grid:
    groups["icp_pairs_kernel_grid"] = sum(l["dram"] for _, l in grid) / len(grid)
occ = {short(n): ls for n, ls in per_kernel.items() if "occ_" in n}
if occ:
    # every kernel of one update: mean per launch x launches per update (rays: count + fill, tiles: two launches)
    n_updates = max(1, min(len(ls) for k, ls in occ.items() if k.startswith("occ_tile_scan")) if any(k.startswith("occ_tile_scan") for k in occ) else 1)
    groups["occupancy_update"] = sum(sum(l["dram"] for l in ls) for ls in occ.values()) / n_updates
out = dict(source_hash=bench.source_hash(), command=command, dram_bytes_per_launch=groups, kernels=detail,
           how="ncu --set full --clock-control none; dram__bytes_read.sum + dram__bytes_write.sum per launch")
path = os.path.join(ROOT, "profiles", "r02_ncu_traffic.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print(json.dumps(dict(source_hash=out["source_hash"], groups=groups), indent=1))
