"""Shared synthetic-input builders for the parity tests (seeded, small)."""
import numpy as np

from oracle import binding as ob


def random_cells(rng, h, w, model, known_frac=0.7):
    """dense [h][w][stride] records with plausible, internally consistent states"""
    st = ob.STRIDE[model]
    c = np.zeros((h, w, st))
    known = rng.random((h, w)) < known_frac
    p = rng.random((h, w))
    if model == ob.CELL_LWW:
        c[..., 0] = np.where(known, p, 0.5); c[..., 1] = np.where(known, rng.random((h, w)), 0); c[..., 2] = known
    elif model == ob.CELL_AFFINE:
        c[..., 0] = np.where(known, p, 0.5); c[..., 1] = known
    elif model == ob.CELL_MEAN:
        c[..., 0] = np.where(known, p, 0.5); c[..., 1] = np.where(known, rng.integers(1, 30, (h, w)), 0)
    elif model in (ob.CELL_TBM_CONSISTENT, ob.CELL_TBM_UNKNOWN_EVEN, ob.CELL_CREDIBILIST):
        b = rng.random((h, w, 3)) + 1e-3
        b /= b.sum(-1, keepdims=True)
        u, e, o = b[..., 0], b[..., 1], b[..., 2]
        u = np.where(known, u, 1.0); e = np.where(known, e, 0.0); o = np.where(known, o, 0.0)
        if model == ob.CELL_TBM_CONSISTENT:
            q = o + e
            pp = np.where(known, o / np.where(q == 0, 1, q), 0.5)
            qq = np.where(known, q, 1.0)
        else:
            pp = np.where(known, o + 0.5 * u, 0.5); qq = np.ones((h, w))
        c[..., 0] = pp; c[..., 1] = qq; c[..., 2] = u; c[..., 3] = e; c[..., 4] = o; c[..., 5] = known
    elif model == ob.CELL_GMAPPING:
        tries = np.where(known, rng.integers(1, 20, (h, w)), 0)
        hits = np.where(known, (tries * rng.random((h, w))).astype(int), 0)
        c[..., 0] = np.where(known, hits / np.maximum(tries, 1) * 0.95, -1)
        c[..., 1] = rng.uniform(-3, 3, (h, w)) * (hits > 0); c[..., 2] = rng.uniform(-3, 3, (h, w)) * (hits > 0)
        c[..., 3] = hits; c[..., 4] = tries
    return c


def room_scan(rng, n, fov, half_w=4.0, half_h=3.0, pose=(0.0, 0.0, 0.0), noise=0.0):
    """ranges of a rectangular room seen from `pose` (a closed world: every beam hits)"""
    ang = np.linspace(-fov / 2, fov / 2, n, endpoint=False) if fov < 2 * np.pi - 1e-9 else \
        -np.pi + 2 * np.pi * np.arange(n) / n
    th = ang + pose[2]
    c, s = np.cos(th), np.sin(th)
    with np.errstate(divide="ignore"):
        tx = np.where(c > 0, (half_w - pose[0]) / c, np.where(c < 0, (-half_w - pose[0]) / c, np.inf))
        ty = np.where(s > 0, (half_h - pose[1]) / s, np.where(s < 0, (-half_h - pose[1]) / s, np.inf))
    r = np.minimum(tx, ty)
    if noise:
        r = r + rng.normal(0, noise, n)
    return r, ang


def room_map_cells(rng, h, w, scale, model, half_w=4.0, half_h=3.0, passes=1):
    """a map holding the same room: occupied ring, free interior, unknown outside"""
    om = ob.OracleMap(w, h, scale, model, ob.GROW_NONE)
    est = ob.estimator(ob.EST_CONST)
    for k in range(passes):
        pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-np.pi, np.pi))
        r, a = room_scan(rng, 720, 2 * np.pi, half_w, half_h, pose)
        sc = ob.OracleScan(r, a)
        om.append_scan(sc, pose, 1.0, 0, est, blur=0.3)
    return om.cells().copy()
