import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_constructor_b200 as sg
ctx = sg.Context(0)
rng = np.random.default_rng(1)
for name, size, scale, beams, fov, model, est_kind in (("tiny", 100, 0.1, 360, 2 * np.pi, sg.CELL_MEAN, sg.EST_CONST),
                                                     ("viny", 800, 0.05, 1081, 1.5 * np.pi, sg.CELL_TBM_CONSISTENT, sg.EST_AREA)):
    hw, hh = size * scale * 0.35, size * scale * 0.3
    gm = sg.GridMap(ctx, size, size, scale, model, sg.GROW_PLAIN)
    pose = np.array([0.317, -0.223, 0.1])
    r, a = bench.room_ranges(rng, beams, fov, hw, hh, pose, 0.01)
    scan = sg.Scan(ctx, r, a)
    tbm = model == sg.CELL_TBM_CONSISTENT
    est = sg.estimator(est_kind, occ=(0.95, 0.04) if tbm else (0.95, 1.0), empty=(0.01, 0.003) if tbm else (0.01, 1.0), shift=0.01 * scale)
    for _ in range(3):
        ctx.append_scan(gm, scan, pose, 0.9, 0, est, blur=0.3)
