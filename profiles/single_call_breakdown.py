#!/usr/bin/env python
"""Device-side split of ONE registration call (the library's CUDA events) next to its wall time: a C2 scan pair and the
teapot pair (C1, 3-D point-to-point)."""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth  # noqa: E402

api.init(0)
scans, poses = synth.make_sequence(40, world="room", seed=0)
cfg2 = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
cases = [("C2 scan pair", scans[10], scans[11], cfg2)]
golden = os.path.join(ROOT, "tests", "golden", "teapot.npz")
if os.path.exists(golden):                                      # C1: demos/teapot_icp_demo.py:58-65
    g = np.load(golden)
    cases.append(("C1 teapot", g["moved"], g["teapot"], dict(error_threshold=1e-12, max_iterations=300, voxel_size=0.005, method="point_to_point")))
for name, s, t, cfg in cases:
    for _ in range(5):
        api.icp_batch([s], [t], **cfg)
    wall = []
    for _ in range(50):
        t0 = time.perf_counter()
        out = api.icp_batch([s], [t], **cfg)
        wall.append(time.perf_counter() - t0)
    st = api.icp_last_stats()
    ph = api.icp_phase_profile()
    print(f"{name}: wall median {np.median(wall) * 1e6:.0f} us | voxel {st['voxel_kernel_ns'] / 1e3:.0f} normals {st['normals_kernel_ns'] / 1e3:.0f} "
          f"pair kernel {st['pair_kernel_ns'] / 1e3:.0f} us | iterations {st['iterations']} swept {st['points_swept']} evals {st['sweep_pair_evals'] / 1e6:.1f} M | "
          f"phases(it>=8) {{{', '.join(f'{k}: {v:.0f}' for k, v in ph['phases'].items())}}}")
