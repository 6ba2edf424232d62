"""Pins the plain-C oracle (oracle/*.c) against the UNMODIFIED reference compiled in
place (oracle/_ref/libslamref.so) on seeded inputs.  CPU only.  Skipped when the
reference build is absent (clean checkout without /root/reference); the committed
fixtures of tests/test_golden.py then carry the pin.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import random_cells, room_map_cells, room_scan
from oracle import binding as ob

ALL_MODELS = [ob.CELL_LWW, ob.CELL_AFFINE, ob.CELL_MEAN, ob.CELL_TBM_CONSISTENT, ob.CELL_TBM_UNKNOWN_EVEN,
              ob.CELL_GMAPPING, ob.CELL_CREDIBILIST]


def test_raycast_matches_reference(refso):
    rng = np.random.default_rng(1)
    for it in range(6000):
        sc = [0.05, 0.1, 0.025, 1.0][it % 4]
        b = rng.uniform(-5, 5, 2)
        e = b + rng.uniform(-6, 6, 2) * (0.02 if it % 7 == 0 else 1)
        if it % 11 == 0: e[0] = b[0]
        if it % 13 == 0: e[1] = b[1]
        if it % 17 == 0: b = np.round(b / sc) * sc
        if it % 19 == 0: e = np.round(e / sc) * sc
        if it % 23 == 0:
            b = (np.floor(b / sc) + 0.5) * sc
            e = (np.floor(e / sc) + 0.5) * sc
        a = ob.raycast(ob.orc.orc_raycast, b[0], b[1], e[0], e[1], sc)
        r = ob.raycast(refso.ref_raycast, b[0], b[1], e[0], e[1], sc)
        assert a.shape == r.shape and (a == r).all(), (b, e, sc)


def test_bresenham_matches_reference(refso):
    rng = np.random.default_rng(2)
    cap = 256
    A = np.zeros((cap, 2), np.int32)
    B = np.zeros((cap, 2), np.int32)
    for _ in range(3000):
        p = [int(v) for v in rng.integers(-40, 40, 4)]
        n1 = ob.orc.orc_bresenham(*p, ob.iptr(A), cap)
        n2 = refso.ref_bresenham(*p, ob.iptr(B), cap)
        assert n1 == n2 and (A[:n1] == B[:n2]).all(), p


def test_rasterize_and_overlap_match_reference(refso):
    rng = np.random.default_rng(3)
    cap = 4096
    B = np.zeros((cap, 2), np.int32)
    for it in range(3000):
        sc = [0.05, 0.1, 0.5][it % 3]
        c = rng.uniform(-3, 3, 2)
        hv, hh = rng.uniform(0, 0.3, 2)
        if it % 5 == 0: hv = hh = 0.0
        if it % 7 == 0: hv = 0.0
        if it % 9 == 0:
            c = np.round(c / sc) * sc
            hv = hh = sc
        bot, top, left, right = c[1] - hv, c[1] + hv, c[0] - hh, c[0] + hh
        for border in (1, 0):
            lbrt = np.zeros(4, np.int32)
            n1 = ob.orc.orc_rasterize_rect(sc, 100, 100, 50, 50, bot, top, left, right, border, ob.iptr(lbrt))
            n2 = refso.ref_rasterize_rect(sc, 100, 100, bot, top, left, right, border, ob.iptr(B), cap)
            assert n1 == n2
            mine = [(x, y) for x in range(lbrt[0], lbrt[2] + 1) for y in range(lbrt[1], lbrt[3] + 1)] if n1 else []
            assert mine == [tuple(v) for v in B[:n2]]
        # overlap of the window with the cell under its centre and a neighbour
        cx, cy = int(np.floor(c[0] / sc)), int(np.floor(c[1] / sc))
        for dx, dy in ((0, 0), (1, 0), (0, -1)):
            cell = (sc * (cy + dy), sc * (cy + dy + 1), sc * (cx + dx), sc * (cx + dx + 1))
            o1 = ob.orc.orc_rect_overlap(bot, top, left, right, *cell)
            o2 = refso.ref_rect_overlap(bot, top, left, right, *cell)
            assert o1 == o2 or o2 == -1  # -1: the reference asserts on this input


def test_area_estimator_matches_reference(refso):
    rng = np.random.default_rng(4)
    sc = 0.05
    est = ob.estimator(ob.EST_AREA, shift=0.01 * 0.05)
    o1, o2 = np.zeros(2), np.zeros(2)
    checked = 0
    for it in range(30000):
        cx, cy = rng.integers(-5, 5, 2)
        cb, ct, cl, cr = sc * cy, sc * (cy + 1), sc * cx, sc * (cx + 1)
        b = rng.uniform(-0.4, 0.4, 2)
        e = rng.uniform(-0.4, 0.4, 2)
        m = it % 10
        if m == 0: e = np.array([rng.uniform(cl, cr), rng.uniform(cb, ct)])
        if m == 1: b = np.array([rng.uniform(cl, cr), rng.uniform(cb, ct)])
        if m == 2: e = np.array([cl, cb]) if it % 20 < 10 else np.array([cr, ct])
        if m == 3: b[1] = e[1] = cb if it % 20 < 10 else ct
        if m == 4: b[0] = e[0] = cl if it % 20 < 10 else cr
        if m == 5: e = np.array([cl, rng.uniform(cb, ct)])
        if m == 6:
            b = np.round(b / sc) * sc
            e = np.round(e / sc) * sc
        if m == 7: e = np.array([rng.uniform(cl, cr), ct])
        for occ in (0, 1):
            ob.orc.orc_estimate_occupancy(C.byref(est), b[0], b[1], e[0], e[1], cb, ct, cl, cr, occ, ob.dptr(o1))
            refso.ref_estimate_occupancy(C.byref(est), b[0], b[1], e[0], e[1], cb, ct, cl, cr, occ, ob.dptr(o2))
            if np.isinf(o2).any():
                continue  # the reference asserts on this input
            checked += 1
            assert np.array_equal(o1, o2, equal_nan=True), (m, occ, b, e)
    assert checked > 50000


@pytest.mark.parametrize("model", ALL_MODELS)
def test_cell_update_and_impact_match_reference(refso, model):
    rng = np.random.default_rng(10 + model)
    cells = random_cells(rng, 40, 40, model).reshape(-1, ob.STRIDE[model])
    for i, rec in enumerate(cells):
        a, b = rec.copy(), rec.copy()
        for k in range(4):
            p, q, quality = rng.random(3)
            if (i + k) % 13 == 0: p = np.nan
            if (i + k) % 5 == 0: p *= 0.4
            obx, oby = rng.uniform(-3, 3, 2)
            for oie in (ob.OIE_DISCREPANCY, ob.OIE_OCCUPANCY):
                v1 = ob.orc.orc_cell_impact(model, oie, ob.dptr(a), obx, oby)
                v2 = refso.ref_cell_impact(model, oie, ob.dptr(b), obx, oby)
                assert v1 == v2 or (np.isnan(v1) and np.isnan(v2))
            ob.orc.orc_cell_update(model, ob.dptr(a), 1, p, q, obx, oby, quality)
            refso.ref_cell_update(model, ob.dptr(b), 1, p, q, obx, oby, quality)
            assert np.array_equal(a, b, equal_nan=True), (model, rec, p, q, quality)


def _both_maps(model, cells, scale, grow=ob.GROW_PLAIN, pyramid_oie=-1):
    h, w = cells.shape[:2]
    om = ob.OracleMap(w, h, scale, model, grow)
    om.set_cells(cells)
    rm = ob.RefMap(w, h, scale, model, grow, pyramid_oie=pyramid_oie)
    rm.set_cells(cells)
    return om, rm


def _ref_scores(refso, rm, r, a, params, spw, f0, poses, occ=None, factor=None, cartesian=0, skip=0, max_range=-1.0):
    poses = ob.f64(poses).reshape(-1, 3)
    out = np.empty(len(poses))
    nf = C.c_int32()
    r, a = ob.f64(r), ob.f64(a)
    occ = np.ascontiguousarray(occ if occ is not None else np.ones(len(r)), dtype=np.uint8)
    refso.ref_score_poses(rm.h_, len(r), ob.dptr(r), ob.dptr(a), ob.u8ptr(occ),
                          ob.dptr(ob.f64(factor)) if factor is not None else None, cartesian, spw, skip, max_range,
                          C.byref(params), f0[0], f0[1], f0[2], ob.dptr(poses), len(poses), ob.dptr(out), C.byref(nf))
    return out, nf.value


def _orc_scores(om, r, a, params, spw, f0, poses, occ=None, skip=0, max_range=-1.0):
    r, a = ob.f64(r), ob.f64(a)
    occ = np.ascontiguousarray(occ if occ is not None else np.ones(len(r)), dtype=np.uint8)
    keep = np.zeros(len(r), np.int32)
    k = ob.orc.orc_filter_scan(om.h_, len(r), ob.dptr(r), ob.dptr(a), ob.u8ptr(occ), f0[0], f0[1], f0[2], skip,
                               max_range, ob.iptr(keep))
    keep = keep[:k]
    fr, fa = r[keep].copy(), a[keep].copy()
    w = np.empty(k)
    ob.orc.orc_point_weights(spw, k, ob.dptr(fr), ob.dptr(fa), ob.dptr(w))
    scan = ob.OracleScan(fr, fa, weight=w)
    return om.score(scan, params, poses), k


@pytest.mark.parametrize("model,oie", [(ob.CELL_LWW, 0), (ob.CELL_MEAN, 0), (ob.CELL_TBM_CONSISTENT, 0),
                                       (ob.CELL_TBM_UNKNOWN_EVEN, 1), (ob.CELL_AFFINE, 1), (ob.CELL_CREDIBILIST, 0)])
@pytest.mark.parametrize("spw", [ob.SPW_EVEN, ob.SPW_VINY, ob.SPW_AHR])
def test_scores_obstacle_mode_match_reference(refso, model, oie, spw):
    rng = np.random.default_rng(100 + model * 7 + spw)
    cells = room_map_cells(rng, 200, 200, 0.05, model, passes=3)
    om, rm = _both_maps(model, cells, 0.05)
    r, a = room_scan(rng, 181, np.deg2rad(270), pose=(0.2, -0.1, 0.3), noise=0.01)
    poses = np.array([0.2, -0.1, 0.3]) + rng.normal(0, [0.2, 0.2, 0.1], (300, 3))
    params = ob.spe_params(ob.OOPE_OBSTACLE, oie)
    s1, k1 = _orc_scores(om, r, a, params, spw, (0.2, -0.1, 0.3), poses)
    s2, k2 = _ref_scores(refso, rm, r, a, params, spw, (0.2, -0.1, 0.3), poses)
    assert k1 == k2 == 181
    assert np.array_equal(s1, s2)  # bit-exact, same libm


@pytest.mark.parametrize("oope", [ob.OOPE_MAX, ob.OOPE_MEAN, ob.OOPE_OVERLAP])
def test_scores_window_modes_match_reference(refso, oope):
    rng = np.random.default_rng(200 + oope)
    cells = room_map_cells(rng, 200, 200, 0.05, ob.CELL_MEAN, passes=3)
    om, rm = _both_maps(ob.CELL_MEAN, cells, 0.05)
    r, a = room_scan(rng, 90, np.deg2rad(240), pose=(0.0, 0.0, 0.0), noise=0.01)
    poses = rng.normal(0, [0.2, 0.2, 0.1], (60, 3))
    for win in ((0.1, 0.1), (0.05, 0.2), (0.0, 0.0), (0.3, 0.0)):
        params = ob.spe_params(oope, ob.OIE_DISCREPANCY, win_v=win[0], win_h=win[1])
        s1, _ = _orc_scores(om, r, a, params, ob.SPW_EVEN, (0, 0, 0), poses)
        s2, _ = _ref_scores(refso, rm, r, a, params, ob.SPW_EVEN, (0, 0, 0), poses)
        assert np.array_equal(s1, s2), win


def test_filter_scan_matches_reference(refso):
    rng = np.random.default_rng(31)
    cells = random_cells(rng, 60, 60, ob.CELL_LWW)
    for grow in (ob.GROW_NONE, ob.GROW_PLAIN):
        om, rm = _both_maps(ob.CELL_LWW, cells, 0.1, grow)
        r = rng.uniform(0.2, 6, 200)
        a = np.linspace(-2, 2, 200)
        occ = (rng.random(200) < 0.8).astype(np.uint8)
        for skip, mr in ((0, -1.0), (3, -1.0), (0, 3.0), (2, 2.5)):
            k1 = np.zeros(200, np.int32)
            k2 = np.zeros(200, np.int32)
            n1 = ob.orc.orc_filter_scan(om.h_, 200, ob.dptr(r), ob.dptr(a), ob.u8ptr(occ), 0.3, 0.2, 0.5, skip, mr, ob.iptr(k1))
            n2 = refso.ref_filter_scan(rm.h_, 200, ob.dptr(r), ob.dptr(a), ob.u8ptr(occ), 0.3, 0.2, 0.5, skip, mr, ob.iptr(k2))
            assert n1 == n2 and (k1[:n1] == k2[:n2]).all()


def test_gmapping_oope_with_cache_matches_reference(refso):
    rng = np.random.default_rng(41)
    cells = room_map_cells(rng, 200, 200, 0.05, ob.CELL_GMAPPING, passes=4)
    om, rm = _both_maps(ob.CELL_GMAPPING, cells, 0.05)
    r, a = room_scan(rng, 720, 2 * np.pi, pose=(0.1, 0.1, 0.0), noise=0.005)
    X = 0.1 + r * np.cos(a)
    Y = 0.1 + r * np.sin(a)
    out2 = np.empty(720)
    refso.ref_gmapping_point_probs(rm.h_, 0.1, 1, 720, ob.dptr(X), ob.dptr(Y), ob.dptr(out2))
    params = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    cache = ob.GmCache(0, 0, -1.0)
    out1 = np.array([ob.orc.orc_point_probability(om.h_, C.byref(params), X[i], Y[i], C.byref(cache)) for i in range(720)])
    assert np.array_equal(out1, out2)
    assert (out1 > 0).sum() > 300


@pytest.mark.parametrize("model,est_type,blur,grow", [
    (ob.CELL_MEAN, ob.EST_CONST, 0.5, ob.GROW_NONE),
    (ob.CELL_TBM_CONSISTENT, ob.EST_AREA, 0.3, ob.GROW_PLAIN),
    (ob.CELL_TBM_UNKNOWN_EVEN, ob.EST_CONST, 0.0, ob.GROW_PLAIN),
    (ob.CELL_AFFINE, ob.EST_AREA, 0.2, ob.GROW_NONE),
    (ob.CELL_LWW, ob.EST_AREA, -0.01, ob.GROW_PLAIN),
    (ob.CELL_GMAPPING, ob.EST_CONST, 0.0, ob.GROW_TILED),
    (ob.CELL_CREDIBILIST, ob.EST_AREA, 0.3, ob.GROW_PLAIN),
])
def test_append_scan_matches_reference(refso, model, est_type, blur, grow):
    rng = np.random.default_rng(300 + model)
    w = h = 120 if grow != ob.GROW_NONE else 240
    om = ob.OracleMap(w, h, 0.05, model, grow)
    rm = ob.RefMap(w, h, 0.05, model, grow)
    est = ob.estimator(est_type, occ=(0.95, 0.04) if model in (3, 4) else (0.95, 1.0),
                       empty=(0.01, 0.003) if model in (3, 4) else (0.01, 1.0), shift=0.01 * 0.05)
    for k in range(3):
        pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3))
        r, a = room_scan(rng, 241, np.deg2rad(270), pose=pose, noise=0.01)
        occ = (rng.random(241) < 0.95).astype(np.uint8)
        sc = ob.OracleScan(r, a, occ=occ)
        omqe = ob.OMQE_AHR if k == 2 else ob.OMQE_IDLE
        om.append_scan(sc, pose, 0.9, 2 if k == 1 else 0, est, blur=blur, max_range=5.5, omqe=omqe)
        refso.ref_append_scan(rm.h_, 241, ob.dptr(sc.a), ob.dptr(sc.b), ob.u8ptr(sc.occ), pose[0], pose[1], pose[2], 0.9,
                              2 if k == 1 else 0, C.byref(est), blur, 5.5, omqe)
        i1, i2 = om.info(), rm.info()
        assert i1 == i2
        assert np.array_equal(om.cells(), rm.export(), equal_nan=True)
    assert (om.cells()[..., 0] != om.cells()[0, 0, 0]).sum() > 1000


@pytest.mark.parametrize("grow,dims", [(ob.GROW_PLAIN, (64, 64)), (ob.GROW_PLAIN, (50, 37)), (ob.GROW_NONE, (128, 128)),
                                       (ob.GROW_TILED, (100, 100))])
def test_pyramid_matches_reference(refso, grow, dims):
    rng = np.random.default_rng(400 + dims[0])
    w, h = dims
    model = ob.CELL_MEAN
    p = ob.orc.orc_pyramid_create(w, h, 0.1, model, grow, None, ob.OIE_DISCREPANCY)
    rm = ob.RefMap(w, h, 0.1, model, grow, pyramid_oie=ob.OIE_DISCREPANCY)
    try:
        assert ob.orc.orc_pyramid_levels(p) == rm.levels()
        span = 0.45 if grow == ob.GROW_NONE else 0.8
        for k in range(1500):
            x = int(rng.integers(-w * span, w * span))
            y = int(rng.integers(-h * span, h * span))
            pr, q = rng.random(), rng.random()
            ob.orc.orc_pyramid_update(p, x, y, 1, pr, 1.0, 0.0, 0.0, q)
            refso.ref_map_update(rm.h_, x, y, 1, pr, 1.0, 0.0, 0.0, q)
        n = ob.orc.orc_pyramid_levels(p)
        assert n == rm.levels()
        for lv in range(n):
            om = ob.OracleMap(model=model, handle=ob.orc.orc_pyramid_level(p, lv), owner=False)
            assert om.info() == rm.info(lv), lv
            assert np.array_equal(om.cells(), rm.export(lv)), lv
        for t in (0.0, 0.05, 0.1, 0.15, 0.4, 1.0, 7.0, 1e9):
            assert ob.orc.orc_pyramid_rescale(p, t) == refso.ref_pyramid_rescale(rm.h_, t)
    finally:
        ob.orc.orc_pyramid_destroy(p)


def test_match_bound_matches_reference(refso):
    rng = np.random.default_rng(500)
    w = h = 256
    model = ob.CELL_MEAN
    p = ob.orc.orc_pyramid_create(w, h, 0.05, model, ob.GROW_PLAIN, None, ob.OIE_DISCREPANCY)
    rm = ob.RefMap(w, h, 0.05, model, ob.GROW_PLAIN, pyramid_oie=ob.OIE_DISCREPANCY)
    est = ob.estimator(ob.EST_CONST)
    try:
        for k in range(3):
            pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3))
            r, a = room_scan(rng, 200, 2 * np.pi, pose=pose)
            sc = ob.OracleScan(r, a)
            ob.orc.orc_pyramid_append_scan(p, C.byref(sc.s), pose[0], pose[1], pose[2], 1.0, 0, C.byref(est), 0.3, np.inf, 0)
            refso.ref_append_scan(rm.h_, 200, ob.dptr(sc.a), ob.dptr(sc.b), ob.u8ptr(sc.occ), pose[0], pose[1], pose[2], 1.0, 0,
                                  C.byref(est), 0.3, np.inf, 0)
        pose = (0.3, -0.2, 0.1)
        r, a = room_scan(rng, 150, np.deg2rad(270), pose=pose)
        params = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY, prerotated=1)
        for rot in (0.0, 0.02, -0.05):
            x, y = r * np.cos(a + rot + pose[2]), r * np.sin(a + rot + pose[2])
            sc = ob.OracleScan(x, y, cartesian=True)
            for win in ((-1, 1, -1, 1), (0, 1, -1, 0), (0.25, 0.5, -0.5, -0.25), (0.05, 0.1, 0.05, 0.1), (0.1, 0.1, -0.2, -0.2),
                        (-0.5, 0.5, 0.0, 0.0)):
                b1 = ob.orc.orc_match_bound(p, C.byref(sc.s), C.byref(params), pose[0], pose[1], pose[2], rot, *win)
                b2 = refso.ref_match_bound(rm.h_, 150, ob.dptr(sc.a), ob.dptr(sc.b), 1, ob.SPW_EVEN, C.byref(params),
                                           pose[0], pose[1], pose[2], rot, *win)
                assert b1 == b2, (rot, win)
    finally:
        ob.orc.orc_pyramid_destroy(p)


def test_bf_enumerator_matches_reference(refso):
    for args in ((-0.5, 0.5, 0.1, -0.5, 0.5, 0.1, -np.deg2rad(5), np.deg2rad(5), np.deg2rad(1)),
                 (-0.5, 0.5, 0.05, -0.5, 0.5, 0.05, -np.deg2rad(10), np.deg2rad(10), np.deg2rad(1)),
                 (-1, 1, 0.02, -1, 1, 0.02, -0.5, 0.5, 0.01), (0, 0, 0.1, -0.2, 0.2, 0.1, 0, 0, 1)):
        base = (0.31, -0.77, 0.123)
        n2 = refso.ref_bf_enumerate(*base, *args, None, 0)
        P2 = np.empty((n2, 3))
        refso.ref_bf_enumerate(*base, *args, ob.dptr(P2), n2)
        P1 = np.empty((n2 + 8, 3))
        xs, ys, ts = np.empty(4096), np.empty(4096), np.empty(4096)
        nx, ny, nt = C.c_int32(), C.c_int32(), C.c_int32()
        n1 = ob.orc.orc_bf_enumerate(*base, *args, ob.dptr(P1), n2 + 8, ob.dptr(xs), nx, ob.dptr(ys), ny, ob.dptr(ts), nt)
        assert n1 == n2 == nx.value * ny.value * nt.value
        assert np.array_equal(P1[:n1], P2)
        grid = np.stack(np.meshgrid(ts[:nt.value], ys[:ny.value], xs[:nx.value], indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
        assert np.array_equal(grid, P2)
    assert n2 == 5  # the degenerate last case: a single x and theta


def _match_inputs(rng, model=ob.CELL_MEAN, n=120):
    cells = room_map_cells(rng, 200, 200, 0.05, model, passes=4)
    om, rm = _both_maps(model, cells, 0.05)
    true_pose = np.array([0.3, -0.2, 0.2])
    r, a = room_scan(rng, n, np.deg2rad(270), pose=true_pose, noise=0.003)
    return om, rm, r, a, true_pose


def test_hill_climbing_matches_reference(refso):
    rng = np.random.default_rng(600)
    om, rm, r, a, tp = _match_inputs(rng)
    params = ob.spe_params()
    occ = np.ones(len(r), np.uint8)
    for k in range(4):
        init = tp + rng.normal(0, [0.08, 0.08, 0.04])
        w = np.full(len(r), 1.0 / len(r))
        sc = ob.OracleScan(r, a, weight=w)
        m1, m2 = ob.MatchResult(), ob.MatchResult()
        ob.orc.orc_match_hill_climbing(om.h_, C.byref(sc.s), C.byref(params), *init, 6, 0.1, 0.1, C.byref(m1), None)
        refso.ref_match_hc(rm.h_, len(r), ob.dptr(sc.a), ob.dptr(sc.b), ob.u8ptr(occ), ob.SPW_EVEN, C.byref(params), *init,
                           6, 0.1, 0.1, C.byref(m2), None, 0)
        assert (m1.best_prob, m1.dx, m1.dy, m1.dth, m1.poses_tested) == (m2.best_prob, m2.dx, m2.dy, m2.dth, m2.poses_tested)
        assert m1.poses_tested > 30


def test_monte_carlo_matches_reference(refso):
    rng = np.random.default_rng(700)
    om, rm, r, a, tp = _match_inputs(rng)
    params = ob.spe_params()
    occ = np.ones(len(r), np.uint8)
    sc = ob.OracleScan(r, a, weight=np.full(len(r), 1.0 / len(r)))
    for seed in (42, 7, 123456):
        init = tp + rng.normal(0, [0.1, 0.1, 0.05])
        m1, m2 = ob.MatchResult(), ob.MatchResult()
        ob.orc.orc_match_monte_carlo(om.h_, C.byref(sc.s), C.byref(params), *init, seed, 0.2, 0.1, 20, 100, C.byref(m1))
        refso.ref_match_mc(rm.h_, len(r), ob.dptr(sc.a), ob.dptr(sc.b), ob.u8ptr(occ), ob.SPW_EVEN, C.byref(params), *init,
                           seed, 0.2, 0.1, 20, 100, C.byref(m2), None, 0)
        assert (m1.best_prob, m1.dx, m1.dy, m1.dth, m1.poses_tested) == (m2.best_prob, m2.dx, m2.dy, m2.dth, m2.poses_tested)


def test_normal_distribution_restatement(refso):
    for seed, mean, sd in ((42, 0.0, 0.2), (1, 1.5, 0.1), (99, 0.0, 1.0)):
        out2 = np.empty(501)
        refso.ref_normal_samples(seed, mean, sd, 501, ob.dptr(out2))

        g, d = ob.Mt19937(), ob.Normal(mean, sd, 0.0, 0)
        ob.orc.orc_mt_seed(C.byref(g), seed)
        ob.orc.orc_normal_sample.restype = C.c_double
        out1 = np.array([ob.orc.orc_normal_sample(C.byref(d), C.byref(g)) for _ in range(501)])
        assert np.array_equal(out1, out2)


def test_brute_force_matches_reference(refso):
    rng = np.random.default_rng(800)
    om, rm, r, a, tp = _match_inputs(rng, n=60)
    params = ob.spe_params()
    occ = np.ones(len(r), np.uint8)
    sc = ob.OracleScan(r, a, weight=np.full(len(r), 1.0 / len(r)))
    init = tp + np.array([0.07, -0.04, 0.02])
    args = (-0.2, 0.2, 0.05, -0.2, 0.2, 0.05, -0.05, 0.05, 0.01)
    n = ob.orc.orc_bf_enumerate(*init, *args, None, 0, None, None, None, None, None, None)
    P = np.empty((n, 3))
    ob.orc.orc_bf_enumerate(*init, *args, ob.dptr(P), n, None, None, None, None, None, None)
    m1, m2 = ob.MatchResult(), ob.MatchResult()
    ob.orc.orc_match_list(om.h_, C.byref(sc.s), C.byref(params), *init, ob.dptr(P), n, C.byref(m1))
    scores2 = np.empty(n + 1)
    refso.ref_match_bf(rm.h_, len(r), ob.dptr(sc.a), ob.dptr(sc.b), ob.u8ptr(occ), ob.SPW_EVEN, C.byref(params), *init,
                       *args, C.byref(m2), ob.dptr(scores2), n + 1)
    assert (m1.best_prob, m1.dx, m1.dy, m1.dth, m1.poses_tested) == (m2.best_prob, m2.dx, m2.dy, m2.dth, m2.poses_tested)
    scores1 = om.score(sc, params, P)
    assert np.array_equal(scores1, scores2[1:])
    best = C.c_double()
    idx = ob.orc.orc_argbest(ob.dptr(scores1), n, scores2[0], C.byref(best))
    assert best.value == m2.best_prob and idx >= 0
    assert np.allclose(P[idx] - init, [m2.dx, m2.dy, m2.dth], atol=0, rtol=0)
