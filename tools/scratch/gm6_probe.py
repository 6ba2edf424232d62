import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench
import slam_constructor_b200 as sg
from tools.latency_bench import timeit
ctx = sg.Context(0)
rng = np.random.default_rng(7)
n, size, scale, beams = 6, 240, 0.05, 360
parts = sg.Particles(ctx, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
est = sg.estimator(sg.EST_CONST)
pose = np.array([0.317, -0.223, 0.1])
r, a = bench.room_ranges(rng, beams, 2 * np.pi, 3.0, 2.5, pose, 0.005)
scan = sg.Scan(ctx, r, a)
poses = pose + rng.normal(0, [0.02, 0.02, 0.01], (n, 3))
for _ in range(3):
    parts.append_scan(scan, poses, est=est)
print("append us", timeit(lambda: parts.append_scan(scan, poses, est=est), n=30, warm=3))
for cache in (0, 2):
    params = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=cache)
    t = timeit(lambda: parts.match_hc(scan, params, poses + [0.05, -0.04, 0.02]), n=30, warm=3)
    _, _, tested = parts.match_hc(scan, params, poses + [0.05, -0.04, 0.02])
    print("hc cache", cache, "us", t, "variant", ctx.score_stats()["variant"], "tested", tested.tolist(), "kernel us", ctx.last_kernel_ms() * 1e3)
params = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2, trig=sg.TRIG_HOST)
print("hc cache 2 lock-step us", timeit(lambda: parts.match_hc(scan, params, poses + [0.05, -0.04, 0.02]), n=10, warm=2))
params = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2)
far = pose + rng.normal(0, [1.5, 1.5, 0.6], (n, 3))
act = np.array([1, 0, 1, 1, 0, 1], np.uint8)
for name, kw in (("far", dict()), ("far+inactive", dict(active=act))):
    _, _, tested = parts.match_hc(scan, params, far, **kw)
    print(name, "variant", ctx.score_stats(), "tested", tested.tolist())
import time
for src in ([0, 0, 2, 3, 3, 5], [1, 1, 1, 1, 1, 1], [0, 1, 2, 3, 4, 5], [5, 4, 3, 2, 1, 0]):
    t0 = time.perf_counter(); parts.resample(np.array(src, np.int32)); t1 = time.perf_counter()
    print("resample", src, "ms", (t1 - t0) * 1e3)
    parts.append_scan(scan, poses, est=est)
