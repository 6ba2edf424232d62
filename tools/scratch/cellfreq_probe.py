import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_constructor_b200 as sg
ctx = sg.Context(0)
rng = np.random.default_rng(1)
size, scale, beams, fov = 800, 0.05, 1081, 1.5 * np.pi
gm = sg.GridMap(ctx, size, size, scale, sg.CELL_TBM_CONSISTENT, sg.GROW_PLAIN)
pose = np.array([0.317, -0.223, 0.1])
hw, hh = size * scale * 0.35, size * scale * 0.3
r, a = bench.room_ranges(rng, beams, fov, hw, hh, pose, 0.01)
scan = sg.Scan(ctx, r, a)
offs, cells = ctx.raycast(gm, scan, pose)
print("robot cell", np.floor(pose[:2] / scale))
u, c = np.unique(cells, axis=0, return_counts=True)
o = np.argsort(-c)[:8]
for k in o: print(u[k], c[k])
est = sg.estimator(sg.EST_AREA, occ=(0.95, 0.04), empty=(0.01, 0.003), shift=0.01 * scale)
for _ in range(3):
    ctx.append_scan(gm, scan, pose, 0.9, 0, est, blur=0.3)
    for k in o[:5]:
        print("  cell", u[k], c[k], "record", gm.read_cell(int(u[k][0]), int(u[k][1])))
