"""K4 from-scratch pyramid build at the configs[4] shape: the command ncu wraps / a quick timing (not a bench value)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import slam_constructor_b200 as sg
ctx = sg.Context(0)
size = 4096
rng = np.random.default_rng(3)
cells = np.zeros((size, size, 2)); cells[..., 0] = rng.random((size, size)); cells[..., 1] = 1
gm = sg.GridMap(ctx, size, size, 0.025, sg.CELL_MEAN, sg.GROW_PLAIN)
gm.upload(cells)
pyr = sg.Pyramid(ctx, gm, sg.OIE_DISCREPANCY)
pyr.build()
ts, ks = [], []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 10):
    ctx.flush_l2(); ctx.sync(); ctx.timer_begin(); pyr.build(); ts.append(ctx.timer_end()); ks.append(ctx.last_kernel_ms())
print("build ms", float(np.median(ts)), "first fused pass ms", float(np.median(ks)), "levels", pyr.levels())
