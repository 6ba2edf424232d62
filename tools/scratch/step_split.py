"""Device time of one whole K1 step (trig + index + score + reduce + finalize) for a share of configs[2], next to the scoring
kernel alone: what a rank of N spends outside the scoring kernel (no exchange here)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import bench
import slam_constructor_b200 as sg

ctx = sg.Context(0)
wl = bench.make_workload()
gm = sg.GridMap(ctx, bench.MAP_SIZE, bench.MAP_SIZE, bench.MAP_SCALE, sg.CELL_MEAN)
gm.upload(wl["cells"])
scan = sg.Scan(ctx, wl["r"], wl["a"])
params = sg.spe_params(sg.OOPE_OBSTACLE, sg.OIE_DISCREPANCY, trig=sg.TRIG_DEVICE)
for nt in (100, 50, 25, 12):
    ctx.stage_grid(scan, params, wl["xs"], wl["ys"], wl["ts"][:nt])
    for _ in range(3):
        ctx.score_launch(gm); ctx.sync()
    steps, kern = [], []
    for _ in range(20):
        ctx.flush_l2(); ctx.sync()
        ctx.timer_begin(); ctx.score_launch(gm); steps.append(ctx.timer_end()); kern.append(ctx.last_kernel_ms())
    print("thetas", nt, "variant", ctx.score_stats()["variant"], "step ms %.4f" % np.median(steps), "scoring kernel ms %.4f" % np.median(kern),
          "rest us %.1f" % (1e3 * (np.median(steps) - np.median(kern))))
