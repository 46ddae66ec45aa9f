"""Per-phase cycles of single registrations picked from the C5 batch (the heaviest ones by the per-pair profile)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "iterative-closest-point-avmi_b200"), ROOT]
from icp_b200 import api, synth
scans, poses = synth.make_sequence(2000, world="room", seed=0)
flat, off = synth.pack_ragged(scans)
pairs = synth.loop_closure_pairs(poses, 8192, seed=0, max_dist=3.0).astype(np.int32)
cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, method="point_to_line", normal_k=12)
api.init(0)
for p in [int(a) for a in sys.argv[1:]] or [7584, 347, 971, 4768, 4524, 100]:
    si, ti = pairs[p:p + 1, 0].copy(), pairs[p:p + 1, 1].copy()
    api.icp_pair_profile(0)
    api.icp_pairs(flat, off, si, ti, **cfg)
    out = api.icp_pairs(flat, off, si, ti, **cfg)
    st = api.icp_last_stats(); ph = api.icp_phase_profile(); pp = api.icp_pair_profile(1)
    print(f"pair {p}: iters {out['iters'][0]} err {out['error'][0]:.3f} | {pp[0, 0] / 1.965e3:.0f} us | swept {st['points_swept']} ({st['points_swept'] / max(st['iterations'], 1):.0f}/it) "
          f"evals {st['sweep_pair_evals'] / 1e6:.1f} M ({st['sweep_pair_evals'] / max(st['points_swept'], 1):.0f}/pt) fallbacks {st['fp64_rescans']} | "
          f"phases {{{', '.join(f'{k}: {v:.0f}' for k, v in ph['phases'].items())}}}")
