"""Host wall clock of the pieces of one end-to-end brute-force call (scan upload, stage, launch, fetch) at configs[2]."""
import os, sys, time
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
nt = int(os.environ.get("NT", "100"))
ts = wl["ts"][:nt]
for _ in range(5):
    scan.upload(wl["r"], wl["a"]); ctx.score_grid(gm, scan, params, wl["xs"], wl["ys"], ts, want_scores=False)
T = {"upload": 0, "stage": 0, "launch": 0, "fetch": 0, "whole": 0}
n = 200
for _ in range(n):
    ctx.sync()
    t0 = time.perf_counter(); scan.upload(wl["r"], wl["a"])
    t1 = time.perf_counter(); ctx.stage_grid(scan, params, wl["xs"], wl["ys"], ts)
    t2 = time.perf_counter(); ctx.score_launch(gm)
    t3 = time.perf_counter(); ctx.score_fetch()
    t4 = time.perf_counter()
    T["upload"] += t1 - t0; T["stage"] += t2 - t1; T["launch"] += t3 - t2; T["fetch"] += t4 - t3
for _ in range(n):
    t0 = time.perf_counter()
    scan.upload(wl["r"], wl["a"]); ctx.score_grid(gm, scan, params, wl["xs"], wl["ys"], ts, want_scores=False)
    T["whole"] += time.perf_counter() - t0
print("thetas", nt, {k: round(v / n * 1e6, 1) for k, v in T.items()}, "us (fetch includes waiting for the kernels)")
