import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_constructor_b200 as sg
from tools.latency_bench import timeit
ctx = sg.Context(0)
rng = np.random.default_rng(1)
gm = sg.GridMap(ctx, 160, 160, 0.05, sg.CELL_MEAN, sg.GROW_PLAIN)
pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
pose = np.array([0.117, 0.223, -0.1])
r, a = bench.room_ranges(rng, 181, 2 * np.pi, 3.0, 2.5, pose, 0.01)
scan = sg.Scan(ctx, r, a)
est = sg.estimator(sg.EST_CONST)
for k in range(4):
    pyr.append_scan(scan, pose + [0.01 * k, -0.01 * k, 0.002 * k], 0.9, 0, est, blur=0.3)
l0 = ctx.launch_count()
t = timeit(lambda: pyr.append_scan(scan, pose, 0.9, 0, est, blur=0.3), n=20, warm=3)
print("append_scan us", t, "launches", (ctx.launch_count() - l0) / 23, "levels", pyr.levels())
params = sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1)
init = pose + [0.07, -0.04, 0.02]
_, _, st = pyr.match_m3rsm(r, a, init, params, 0.4, 0.4, np.deg2rad(3), np.deg2rad(0.5), 0.05)
t = timeit(lambda: pyr.match_m3rsm(r, a, init, params, 0.4, 0.4, np.deg2rad(3), np.deg2rad(0.5), 0.05), n=20, warm=3)
print("match us", t, st)
