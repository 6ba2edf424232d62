#!/usr/bin/env python
"""Runs every kernel family once at the BASELINE shapes (for ncu captures): K1 grid + list + small batch,
K2/K3 scan insertion (area estimator, TBM cells), K4 pyramid (incremental + build), K5 via the M3RSM matcher,
K6 particle round."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import slam_constructor_b200 as sg  # noqa: E402


def main():
    ctx = sg.Context(0)
    rng = np.random.default_rng(2)
    # K1: configs[2]
    wl = bench.make_workload()
    gm = sg.GridMap(ctx, bench.MAP_SIZE, bench.MAP_SIZE, bench.MAP_SCALE, sg.CELL_MEAN)
    gm.upload(wl["cells"])
    scan = sg.Scan(ctx, wl["r"], wl["a"])
    ctx.stage_grid(scan, sg.spe_params(), wl["xs"], wl["ys"], wl["ts"])
    for _ in range(3):
        ctx.score_launch(gm)
    ctx.score_fetch()
    P = np.stack(np.meshgrid(wl["ts"][:4], wl["ys"], wl["xs"], indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
    ctx.score_poses(gm, scan, sg.spe_params(), P)           # list kernel (40 804 poses)
    ctx.score_poses(gm, scan, sg.spe_params(), P[:100])     # two-phase small batch
    # a quarter of the candidate set: the share of one rank of four -> k_score_grid5 (cp.async pipeline)
    ctx.stage_grid(scan, sg.spe_params(), wl["xs"], wl["ys"], wl["ts"][:25])
    for _ in range(2):
        ctx.score_launch(gm)
    ctx.score_fetch()
    # the window OOPEs on the grid: window tables (k_build_wlut) + the v1 grid kernel
    ctx.stage_grid(scan, sg.spe_params(sg.OOPE_MAX, win_v=0.1, win_h=0.1), wl["xs"], wl["ys"], wl["ts"])
    for _ in range(2):
        ctx.score_launch(gm)
    ctx.score_fetch()
    gm.close(); scan.close()
    # K2/K3: configs[1] shape
    gm = sg.GridMap(ctx, 800, 800, 0.05, sg.CELL_TBM_CONSISTENT, sg.GROW_PLAIN)
    est = sg.estimator(sg.EST_AREA, occ=(0.95, 0.04), empty=(0.01, 0.003), shift=0.01 * 0.05)
    pose = np.array([0.317, -0.223, 0.1])
    r, a = bench.room_ranges(rng, 1081, 1.5 * np.pi, 14.0, 12.0, pose, 0.01)
    scan = sg.Scan(ctx, r, a)
    for _ in range(3):
        ctx.append_scan(gm, scan, pose, 0.9, 0, est, blur=0.3)
    ctx.match_hc(gm, scan, sg.spe_params(), pose + [0.06, -0.05, 0.03])   # whole hill-climbing match, one launch
    shifts = rng.normal(0, [0.2, 0.2, 0.1], (120, 3))
    ctx.match_mc(gm, scan, sg.spe_params(), pose + [0.06, -0.05, 0.03], shifts, 20, 100)  # Monte-Carlo segment, one launch
    gm.close()
    # K4/K5: pyramid, configs[4]-like at 2048
    gm = sg.GridMap(ctx, 2048, 2048, 0.025, sg.CELL_MEAN, sg.GROW_PLAIN)
    pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
    for k in range(3):
        pyr.append_scan(scan, pose + 0.05 * k, 1.0, 0, sg.estimator(), blur=0.3)
    pyr.build()
    pyr.match_m3rsm(r, a, pose + [0.05, -0.04, 0.01], sg.spe_params(sg.OOPE_MAX, prerotated=1), 1.0, 1.0, np.deg2rad(5),
                    np.deg2rad(0.1), 0.05)
    pyr.close(); gm.close(); scan.close()
    # K6: configs[3] shape, 64 particles here
    n = 64
    parts = sg.Particles(ctx, n, 512, 512, 0.05, sg.CELL_GMAPPING, sg.GROW_TILED)
    r, a = bench.room_ranges(rng, 720, 2 * np.pi, 6.0, 5.0, pose, 0.01)
    scan = sg.Scan(ctx, r, a)
    poses = pose + rng.normal(0, [0.03, 0.03, 0.01], (n, 3))
    parts.append_scan(scan, poses)
    parts.match_hc(scan, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1), poses, 6, 0.1, 0.1)            # one launch
    parts.match_hc(scan, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2), poses, 6, 0.1, 0.1)  # carried cache: lock step
    parts.resample((np.arange(n) // 2 * 2).astype(np.int32))   # copy-on-write: table copies ...
    parts.append_scan(scan, poses)                             # ... and the clones of the tiles this scan writes (k_copy_tiles)
    parts.close(); scan.close(); ctx.close()
    print("exercise: done")


if __name__ == "__main__":
    main()
