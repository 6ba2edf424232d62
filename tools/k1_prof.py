"""K1 at the configs[2] shape, a few launches: the command ncu wraps (tools/..., not a bench: no number printed here counts)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import slam_constructor_b200 as sg

n = int(sys.argv[1]) if len(sys.argv) > 1 else 3
ctx = sg.Context(0)
wl = bench.make_workload()
gm = sg.GridMap(ctx, bench.MAP_SIZE, bench.MAP_SIZE, bench.MAP_SCALE, sg.CELL_MEAN)
gm.upload(wl["cells"])
scan = sg.Scan(ctx, wl["r"], wl["a"])
params = sg.spe_params(sg.OOPE_OBSTACLE, sg.OIE_DISCREPANCY, trig=sg.TRIG_DEVICE)
nt = int(os.environ.get("NT", "0")) or len(wl["ts"])
ctx.stage_grid(scan, params, wl["xs"], wl["ys"], wl["ts"][:nt])
tot = 0.0
for _ in range(n):
    if not os.environ.get("NOFLUSH"):
        ctx.flush_l2()
    ctx.score_launch(gm)
    ctx.sync()
    tot += ctx.last_kernel_ms()
_, idx, best = ctx.score_fetch()
print("nt", nt, "variant", ctx.score_stats()["variant"], "kernel ms (not a bench value)", tot / n, idx, best)
if os.environ.get("WIN"):
    # the same candidate set under the max / mean window OOPEs (0.1 x 0.1 m windows = 3 x 3 cells)
    for code, name in ((sg.OOPE_MAX, "max"), (sg.OOPE_MEAN, "mean")):
        pw = sg.spe_params(code, sg.OIE_DISCREPANCY, win_v=0.1, win_h=0.1, trig=sg.TRIG_DEVICE)
        ctx.stage_grid(scan, pw, wl["xs"], wl["ys"], wl["ts"][:nt])
        tot = 0.0
        for _ in range(n):
            ctx.flush_l2(); ctx.score_launch(gm); ctx.sync(); tot += ctx.last_kernel_ms()
        _, idx, best = ctx.score_fetch()
        print("window OOPE", name, "variant", ctx.score_stats()["variant"], "kernel ms (not a bench value)", tot / n, idx, best,
              "window evals/s", len(wl["xs"]) * len(wl["ys"]) * nt * 1081 / (tot / n * 1e-3))
