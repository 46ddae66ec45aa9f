"""Pin the oracle to the live reference and write the golden fixtures.

Run in the BUILD container only (it imports /root/reference, which does not
exist on the GPU box):

    python oracle/pin_against_reference.py

For every case the reference implementation (``utilities.icp`` /
``utilities.mapping`` imported unmodified from /root/reference, with a stub
``pyvista`` module because mapping.py:2 imports it at module top) and the
oracle restatement are run on the same seeded inputs and must agree
BIT-FOR-BIT; only then are inputs + reference outputs written under
``tests/golden/``.  The fixtures are what the GPU-box tests compare against.
"""
from __future__ import annotations

import contextlib
import io
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF = "/root/reference"

sys.dont_write_bytecode = True
sys.path.insert(0, os.path.join(ROOT, "iterative-closest-point-avmi_b200"))
sys.path.insert(0, ROOT)

from icp_b200 import synth                                     # noqa: E402
from oracle import icp_oracle, occupancy_oracle                # noqa: E402


def load_reference():
    sys.modules.setdefault("pyvista", types.ModuleType("pyvista"))
    # the reference package is also called ``utilities``; import it under a
    # private name so it cannot be confused with this repo's drop-in shim
    import importlib.util
    pkg = types.ModuleType("refutil")
    pkg.__path__ = [os.path.join(REF, "utilities")]
    sys.modules["refutil"] = pkg
    mods = {}
    for name in ("icp", "mapping"):
        spec = importlib.util.spec_from_file_location(
            f"refutil.{name}", os.path.join(REF, "utilities", f"{name}.py"))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[f"refutil.{name}"] = mod
        spec.loader.exec_module(mod)
        mods[name] = mod
    return mods["icp"], mods["mapping"]


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def same_bits(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return a.shape == b.shape and a.dtype == b.dtype and a.tobytes() == b.tobytes()


def check(cond, what):
    if not cond:
        raise SystemExit(f"PIN FAILED: {what}")
    print(f"  ok  {what}")


def pin_icp_case(ref_icp, name, src, tgt, kw):
    r0, t0, e0 = quiet(ref_icp.ICP, src, tgt, **kw)
    trace = {}
    r1, t1, e1, iters, status = icp_oracle.register(src, tgt, trace=trace, **kw)
    check(same_bits(r0, r1) and same_bits(t0, t1) and
          np.float64(e0).tobytes() == np.float64(e1).tobytes(),
          f"ICP {name}: oracle == reference bit-for-bit (iters={iters}, status={status})")
    return dict(R=r0, t=t0, err=np.float64(e0), iters=np.int32(iters),
                status=np.int32(status), n_src=np.int32(len(trace["src"])),
                n_tgt=np.int32(len(trace["tgt"])),
                first_match=trace["matches"][0].astype(np.int32)
                if trace["matches"] else np.zeros(0, np.int32))


def main():
    os.makedirs(GOLDEN, exist_ok=True)
    ref_icp, ref_map = load_reference()

    # ---------------------------------------------------------------- teapot
    print("teapot (demos/teapot_icp_demo.py:28-65)")
    teapot = np.loadtxt(os.path.join(REF, "teapot.csv"), delimiter=",")
    ang = np.radians(25.0)
    ry = np.array([[np.cos(ang), 0, np.sin(ang)], [0, 1, 0], [-np.sin(ang), 0, np.cos(ang)]])
    shift = np.array([0.25, 0.05, 0.0])
    moved = teapot @ ry.T + shift
    kw = dict(error_threshold=1e-12, max_iterations=300, voxel_size=0.005,
              method="point_to_point")
    out = pin_icp_case(ref_icp, "teapot p2p 3D", moved, teapot, kw)
    # known answer: the inverse of the applied transform
    check(np.abs(out["R"] - ry.T).max() < 1e-12 and
          np.abs(out["t"] + ry.T @ shift).max() < 1e-12, "teapot known-answer")
    # point_to_line on 3-D silently means point_to_point (icp.py:162)
    r3, t3, e3 = quiet(ref_icp.ICP, moved, teapot, 1e-12, 300, 0.005, method="point_to_line")
    check(same_bits(r3, out["R"]) and same_bits(t3, out["t"]), "3-D p2l == p2p")
    np.savez_compressed(os.path.join(GOLDEN, "teapot.npz"), teapot=teapot, moved=moved,
                        ry=ry, shift=shift, **out)

    # ------------------------------------------------------- voxel + normals
    print("voxel_downsample / estimate_normals_2d (icp.py:117-129, 51-76)")
    scans, poses = synth.make_sequence(12, world="room", seed=0)
    vox = {}
    for tag, cloud, v in (("scan0_v004", scans[0], 0.04), ("scan1_v006", scans[1], 0.06),
                          ("scan2_v030", scans[2], 0.30), ("teapot_v0005", teapot, 0.005),
                          ("teapot_v05", teapot, 0.5)):
        a = ref_icp.voxel_downsample(cloud, v)
        b = icp_oracle.voxel_means(cloud, v)
        check(same_bits(a, b), f"voxel {tag}: {len(cloud)} -> {len(a)}")
        vox[f"{tag}_in"], vox[f"{tag}_out"], vox[f"{tag}_v"] = cloud, a, np.float64(v)
    np.savez_compressed(os.path.join(GOLDEN, "voxel.npz"), **vox)

    nrm = {}
    for tag, cloud, k in (("scan0_k12", vox["scan0_v004_out"], 12),
                          ("scan1_k10", vox["scan1_v006_out"], 10),
                          ("tiny_k12", vox["scan2_v030_out"][:9], 12)):
        a = ref_icp.estimate_normals_2d(cloud, k=k)
        b, nbr = icp_oracle.pca_normals_2d(cloud, k=k, return_neighbours=True)
        check(same_bits(a, b), f"normals {tag}")
        nrm[f"{tag}_in"], nrm[f"{tag}_out"], nrm[f"{tag}_k"] = cloud, a, np.int32(k)
        nrm[f"{tag}_nbr"] = nbr.astype(np.int32)
    np.savez_compressed(os.path.join(GOLDEN, "normals.npz"), **nrm)

    # ------------------------------------------------------------ 2-D ICP
    print("2-D ICP cases (icp.py:132-223; params config.yaml:19-24, slam.py:92-97, 217-225)")
    cfg = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04,
               method="point_to_line", normal_k=12)
    dflt = dict(error_threshold=1e-7, max_iterations=100, voxel_size=0.06,
                method="point_to_line", normal_k=10)
    cases = {}

    def add(name, src, tgt, kw):
        res = pin_icp_case(ref_icp, name, src, tgt, kw)
        cases[f"{name}/src"], cases[f"{name}/tgt"] = src, tgt
        for key, val in kw.items():
            if val is not None:
                cases[f"{name}/kw_{key}"] = np.asarray(val)
        for key, val in res.items():
            cases[f"{name}/{key}"] = val

    for i in range(5):
        add(f"p2l_cfg_{i}", scans[i], scans[i + 1], cfg)
    add("p2l_default_0", scans[5], scans[6], dflt)
    add("p2l_default_1", scans[8], scans[6], dflt)
    th = 0.02
    rinit = np.array([[np.cos(th), -np.sin(th)], [np.sin(th), np.cos(th)]])
    add("p2l_init", scans[6], scans[7], dict(cfg, R_init=rinit, t_init=np.array([0.03, -0.02])))
    add("p2p_2d", scans[7], scans[8], dict(cfg, method="point_to_point"))
    # scan -> small submap with correspondence gate (slam.py:217-225 shape)
    sub = np.vstack([synth.to_world_frame(scans[i], poses[i]) for i in range(4)])
    sub = ref_icp.voxel_downsample(sub, 0.04)
    x, y, a = poses[4]
    rin = np.array([[np.cos(a + 0.01), -np.sin(a + 0.01)], [np.sin(a + 0.01), np.cos(a + 0.01)]])
    tin = np.array([x + 0.05, y - 0.04])
    gate = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04,
                R_init=rin, t_init=tin, method="point_to_point", max_corr_dist=1.5)
    add("submap_gate", scans[4], sub, gate)
    # gate so tight that iteration 0 breaks: returns (R_init, t_init, inf) (icp.py:186-187)
    add("gate_break", scans[4], sub, dict(gate, max_corr_dist=1e-6))
    # far-off start: runs into max_iterations or a late break
    add("few_iters", scans[0], scans[1], dict(cfg, max_iterations=3))
    np.savez_compressed(os.path.join(GOLDEN, "icp2d.npz"), **cases)

    # ------------------------------------------ scan -> large submap (grid path)
    print("scan -> 20-scan submap, p2p + gate (slam.py:103-108, 217-225)")
    big_scans, big_poses = synth.make_sequence(22, world="room", seed=21, traj_seed=5)
    raw_sub = np.vstack([synth.to_world_frame(big_scans[i], big_poses[i]) for i in range(20)])
    submap = ref_icp.voxel_downsample(raw_sub, 0.04)                 # slam.py:108
    check(same_bits(submap, icp_oracle.voxel_means(raw_sub, 0.04)), f"big voxel {len(raw_sub)} -> {len(submap)}")
    big = {}
    for tag, si, (ex, ey, eth) in (("a", 20, (0.05, -0.04, 0.01)), ("b", 21, (-0.08, 0.06, -0.02))):
        x, y, a = big_poses[si]
        rin = np.array([[np.cos(a + eth), -np.sin(a + eth)], [np.sin(a + eth), np.cos(a + eth)]])
        tin = np.array([x + ex, y + ey])
        kwb = dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, R_init=rin, t_init=tin,
                   method="point_to_point", max_corr_dist=1.5)
        res = pin_icp_case(ref_icp, f"submap_big_{tag}", big_scans[si], submap, kwb)
        big[f"{tag}/src"], big[f"{tag}/R_init"], big[f"{tag}/t_init"] = big_scans[si], rin, tin
        for key, val in res.items():
            big[f"{tag}/{key}"] = val
    # the un-downsampled stack as the target: 21k raw points go through ICP's own voxel pass (icp.py:151)
    resr = pin_icp_case(ref_icp, "submap_raw_target", big_scans[20], raw_sub,
                        dict(error_threshold=1e-10, max_iterations=150, voxel_size=0.04, R_init=big["a/R_init"],
                             t_init=big["a/t_init"], method="point_to_point", max_corr_dist=1.5))
    for key, val in resr.items():
        big[f"raw/{key}"] = val
    resl = pin_icp_case(ref_icp, "submap_big_p2l", big_scans[20], submap,
                        dict(cfg, R_init=big["a/R_init"], t_init=big["a/t_init"]))
    for key, val in resl.items():
        big[f"p2l/{key}"] = val
    np.savez_compressed(os.path.join(GOLDEN, "submap.npz"), raw_sub=raw_sub, submap=submap, **big)

    # ---------------------------------------------------------- bresenham
    print("Bresenham cells (mapping.py:68-89)")
    rng = np.random.default_rng(11)
    ends = rng.integers(-60, 60, size=(400, 4))
    ends[:8] = [[0, 0, 0, 0], [5, 5, 5, 9], [5, 5, 9, 5], [3, 3, 7, 7], [3, 3, -1, 7],
                [0, 0, 10, 5], [0, 0, 5, 10], [2, 2, -9, -4]]
    flat, off = [], [0]
    for x0, y0, x1, y1 in ends:
        cells = ref_map.OccupancyGrid2D._bresenham(int(x0), int(y0), int(x1), int(y1))
        mine = occupancy_oracle.line_cells_py(int(x0), int(y0), int(x1), int(y1))
        c_cells = occupancy_oracle.line_cells_c(int(x0), int(y0), int(x1), int(y1))
        assert cells == mine and [tuple(c) for c in c_cells] == cells
        assert len(cells) == max(abs(x1 - x0), abs(y1 - y0))
        flat.extend(cells)
        off.append(len(flat))
    check(True, f"{len(ends)} rays: reference == py oracle == C oracle, len == max(|dx|,|dy|)")
    np.savez_compressed(os.path.join(GOLDEN, "bresenham.npz"), ends=ends,
                        cells=np.asarray(flat, dtype=np.int32).reshape(-1, 2),
                        off=np.asarray(off, dtype=np.int64))

    # ---------------------------------------------------------- occupancy
    print("OccupancyGrid2D.update_scan (mapping.py:103-141)")
    gkw = dict(resolution=0.05, p_hit=0.85, p_miss=0.42, log_odds_min=-8.0, log_odds_max=8.0)
    bounds = (-9.0, 7.0, -6.0, 8.5)                 # 320 x 290 cells
    wscans, wposes = synth.make_sequence(40, world="room", seed=5, traj_seed=3)
    rngo = np.random.default_rng(2)
    origins, hits = [], []
    for s in range(40):
        pts = synth.to_world_frame(wscans[s], wposes[s]) * 0.35   # shrink so most rays land in-grid
        org = wposes[s, :2] * 0.35
        if s % 7 == 3:                                            # duplicates + zero-length rays
            pts = np.vstack([pts, pts[:50], np.tile(org, (5, 1))])
        if s == 9:
            pts = np.zeros((0, 2))                                # empty scan (mapping.py:113)
        if s == 12:
            org = np.array([-11.0, 0.3])                          # origin outside the grid
        if s == 20:
            pts = pts[rngo.permutation(len(pts))[:200]] * 3.0     # many out-of-bounds endpoints
        origins.append(org)
        hits.append(pts)
    g_ref = ref_map.OccupancyGrid2D(*bounds, **gkw)
    g_py = occupancy_oracle.GridOraclePy(*bounds, **gkw)
    g_c = occupancy_oracle.GridOracleC(*bounds, **gkw)
    g_cf = occupancy_oracle.GridOracleC(*bounds, **gkw)
    snaps = {}
    for s in range(40):
        g_ref.update_scan(origins[s], hits[s])
        g_c.update_scan(origins[s], hits[s])
        g_cf.update_scan(origins[s], hits[s], fast=True)
        if s < 14:
            g_py.update_scan(origins[s], hits[s])
            assert same_bits(g_ref.log_odds, g_py.log_odds), s
        assert same_bits(g_ref.log_odds, g_c.log_odds), s
        assert same_bits(g_ref.log_odds, g_cf.log_odds), s
        if s in (0, 13, 39):
            snaps[f"snap_{s}"] = g_ref.log_odds.copy()
    check(True, "40 scans: reference == py oracle (first 14) == C oracle == C oracle(bbox clip)")
    check(g_ref.l_hit == 1.7346010553881064 and g_ref.l_miss == -0.3227733922630512,
          "log-odds increments (SURVEY §8(a) M1)")
    # a clamp interval that excludes 0: the first update clips every cell (mapping.py:141)
    odd = dict(gkw, log_odds_min=0.5, log_odds_max=3.0)
    o_ref = ref_map.OccupancyGrid2D(*bounds, **odd)
    o_c = occupancy_oracle.GridOracleC(*bounds, **odd)
    for s in range(3):
        o_ref.update_scan(origins[s], hits[s])
        o_c.update_scan(origins[s], hits[s], fast=True)
    check(same_bits(o_ref.log_odds, o_c.log_odds), "clamp interval excluding 0")
    flat_hits, hit_off = synth.pack_ragged(hits)
    np.savez_compressed(os.path.join(GOLDEN, "occupancy.npz"), bounds=np.asarray(bounds),
                        origins=np.asarray(origins), hits=flat_hits, hit_off=hit_off,
                        odd_final=o_ref.log_odds, **{k: np.float64(v) for k, v in gkw.items()},
                        **snaps)
    sizes = {f: os.path.getsize(os.path.join(GOLDEN, f)) for f in sorted(os.listdir(GOLDEN))}
    print("fixtures:", sizes)


if __name__ == "__main__":
    main()
