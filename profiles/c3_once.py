"""One 64-pair scan -> submap batch (C3, hash-grid nearest neighbour) and one single call: the command the ncu
captures of the big-target kernels are taken from (big_voxel_kernel, big_grid_kernel, icp_pairs_kernel<2, grid>)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
import bench
from icp_b200 import api
api.init(0)
target, flat, off, R0, t0s = bench.build_c3()
kw = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_point", max_corr_dist=1.5,
          R_init=np.asarray(R0), t_init=np.asarray(t0s))
si, ti = np.arange(1, 65, dtype=np.int32), np.zeros(64, dtype=np.int32)
for rep in range(3):
    t0 = time.perf_counter()
    out = api.icp_pairs(flat, off, si, ti, **kw)
    dt = time.perf_counter() - t0
ks, ex = api.icp_last_stats(), api.icp_extra_stats()
print(f"64 pairs: {dt * 1e3:.2f} ms; voxel {ks['voxel_kernel_ns'] / 1e6:.3f} grid {ks['normals_kernel_ns'] / 1e6:.3f} pairs {ks['pair_kernel_ns'] / 1e6:.3f} ms; "
      f"mean iterations {out['iters'].mean():.1f}; {ex}")
print(bench.c3_roofline(ks, ex, len(target)))
