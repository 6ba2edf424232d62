import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_constructor_b200 as sg
ctx = sg.Context(0)
n, size, scale, beams = 256, 2560, 0.05, 720
rng = np.random.default_rng(7)
parts = sg.Particles(ctx, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
est = sg.estimator(sg.EST_CONST)
pose = np.array([0.317, -0.223, 0.1])
r, a = bench.room_ranges(rng, beams, 2 * np.pi, 17.5, 15.0, pose, 0.01)
scan = sg.Scan(ctx, r, a)
poses = pose + rng.normal(0, [0.02, 0.02, 0.01], (n, 3))
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    parts.append_scan(scan, poses, est=est)
