"""K4/K5 parity on the GPU: the max-pyramid (incremental maintenance through scan insertion, and the
from-scratch build) and the batched Match upper bounds, against the CPU oracle.  The incremental fold is
order-exact, so level records are asserted bit-equal; bounds are asserted bit-equal too (max + ordered sum)."""
import ctypes as C

import numpy as np
import pytest

from helpers import room_scan
from oracle import binding as ob

pytestmark = pytest.mark.gpu
ATOL = 1e-6


def _compare_levels(pyr, op, model):
    n = ob.orc.orc_pyramid_levels(op)
    assert pyr.levels() == n
    for lv in range(n):
        om = ob.OracleMap(model=model, handle=ob.orc.orc_pyramid_level(op, lv), owner=False)
        assert pyr.level_info(lv) == om.info(), lv
        got, imp = pyr.level(lv, want_impact=True)
        want = om.cells()
        assert np.array_equal(got, want, equal_nan=True), lv
        olut, _ = om.lut(0)
        assert np.allclose(imp, olut, rtol=0, atol=ATOL, equal_nan=True) and np.array_equal(imp, olut, equal_nan=True)


@pytest.mark.parametrize("model,grow,dims,scale", [
    (ob.CELL_MEAN, ob.GROW_PLAIN, (256, 256), 0.05),
    (ob.CELL_TBM_CONSISTENT, ob.GROW_PLAIN, (200, 150), 0.05),
    (ob.CELL_MEAN, ob.GROW_NONE, (256, 256), 0.05),
    (ob.CELL_LWW, ob.GROW_PLAIN, (64, 64), 0.1),      # the room is larger than the map: levels grow and multiply
    (ob.CELL_GMAPPING, ob.GROW_TILED, (100, 100), 0.1),
])
def test_pyramid_incremental_matches_oracle(sg, gpu, model, grow, dims, scale):
    rng = np.random.default_rng(3000 + model + dims[0])
    w, h = dims
    op = ob.orc.orc_pyramid_create(w, h, scale, model, grow, None, ob.OIE_DISCREPANCY)
    gm = sg.GridMap(gpu, w, h, scale, model, grow)
    pyr = sg.Pyramid(gpu, gm, sg.OIE_DISCREPANCY)
    tbm = model in (3, 4)
    kw = dict(occ=(0.95, 0.04) if tbm else (0.95, 1.0), empty=(0.01, 0.003) if tbm else (0.01, 1.0))
    oest, gest = ob.estimator(ob.EST_CONST, **kw), sg.estimator(sg.EST_CONST, **kw)
    try:
        assert pyr.levels() == ob.orc.orc_pyramid_levels(op)
        half = (4.0, 3.0) if grow != ob.GROW_NONE else (min(w, h) * scale * 0.3, min(w, h) * scale * 0.25)
        for k in range(4):
            pose = (rng.uniform(-0.5, 0.5), rng.uniform(-0.5, 0.5), rng.uniform(-3, 3))
            r, a = room_scan(rng, 200, 2 * np.pi, half_w=half[0], half_h=half[1], pose=pose, noise=0.02)
            sc = ob.OracleScan(r, a)
            n1 = ob.orc.orc_pyramid_append_scan(op, C.byref(sc.s), pose[0], pose[1], pose[2], 0.8, 0, C.byref(oest), 0.3, np.inf, 0)
            gsc = sg.Scan(gpu, r, a)
            n2 = pyr.append_scan(gsc, pose, 0.8, 0, gest, blur=0.3)
            gsc.close()
            assert n1 == n2
            _compare_levels(pyr, op, model)
        for t in (0.0, scale / 2, scale, scale * 1.5, scale * 4, 1.0, 7.0, 1e9):
            assert pyr.rescale(t) == ob.orc.orc_pyramid_rescale(op, t)
    finally:
        ob.orc.orc_pyramid_destroy(op)
        pyr.close(); gm.close()


def test_pyramid_build_is_max_of_covered_cells(sg, gpu):
    """the reference's own invariant (m3rsm_rescalable_map_test.cpp verify_map_state): after writing
    distinct values every coarse cell equals the max of the fine cells it covers; and the same values
    written once through the oracle's incremental rule give the same levels"""
    rng = np.random.default_rng(3100)
    for (w, h) in ((64, 64), (50, 37), (128, 96)):
        vals = rng.permutation(w * h).reshape(h, w) / float(w * h)   # distinct, > eps apart
        cells = np.zeros((h, w, 3)); cells[..., 0] = vals; cells[..., 1] = 1; cells[..., 2] = 1
        gm = sg.GridMap(gpu, w, h, 0.1, sg.CELL_LWW, sg.GROW_PLAIN)
        gm.upload(cells)
        pyr = sg.Pyramid(gpu, gm, sg.OIE_OCCUPANCY)
        pyr.build()
        op = ob.orc.orc_pyramid_create(w, h, 0.1, ob.CELL_LWW, ob.GROW_PLAIN, None, ob.OIE_OCCUPANCY)
        for y in range(h):
            for x in range(w):
                ob.orc.orc_pyramid_update(op, x - w // 2, y - h // 2, 1, float(vals[y, x]), 1.0, 0.0, 0.0, 1.0)
        n = ob.orc.orc_pyramid_levels(op)
        assert pyr.levels() == n
        for lv in range(1, n):
            om = ob.OracleMap(model=ob.CELL_LWW, handle=ob.orc.orc_pyramid_level(op, lv), owner=False)
            assert pyr.level_info(lv) == om.info()
            got = pyr.level(lv)
            assert np.array_equal(got, om.cells()), lv
        top = pyr.level(n - 1)
        assert top.shape[:2] == (1, 1) and top[0, 0, 0] == vals.max()
        # direct check of the invariant on level 1 (2x2 blocks in world alignment)
        i1 = pyr.level_info(1)
        l1 = pyr.level(1)[..., 0]
        for Y in range(i1["h"]):
            for X in range(i1["w"]):
                ex, ey = X - i1["ox"], Y - i1["oy"]
                xs = [2 * ex + d + w // 2 for d in (0, 1)]; ys = [2 * ey + d + h // 2 for d in (0, 1)]
                blk = [vals[yy, xx] for yy in ys for xx in xs if 0 <= xx < w and 0 <= yy < h]
                if blk:
                    assert l1[Y, X] == max(blk)
        ob.orc.orc_pyramid_destroy(op)
        pyr.close(); gm.close()


def test_match_bounds_bit_exact(sg, gpu):
    rng = np.random.default_rng(3200)
    w = h = 256
    model = ob.CELL_MEAN
    op = ob.orc.orc_pyramid_create(w, h, 0.05, model, ob.GROW_PLAIN, None, ob.OIE_DISCREPANCY)
    gm = sg.GridMap(gpu, w, h, 0.05, model, sg.GROW_PLAIN)
    pyr = sg.Pyramid(gpu, gm, sg.OIE_DISCREPANCY)
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    try:
        for k in range(3):
            pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3))
            r, a = room_scan(rng, 200, 2 * np.pi, pose=pose)
            sc = ob.OracleScan(r, a)
            ob.orc.orc_pyramid_append_scan(op, C.byref(sc.s), pose[0], pose[1], pose[2], 1.0, 0, C.byref(oest), 0.3, np.inf, 0)
            gsc = sg.Scan(gpu, r, a)
            pyr.append_scan(gsc, pose, 1.0, 0, gest, blur=0.3)
            gsc.close()
        pose = (0.3, -0.2, 0.1)
        r, a = room_scan(rng, 150, np.deg2rad(270), pose=pose)
        oparams = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY, prerotated=1)
        gparams = sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1)
        rots = (0.0, 0.02, -0.05)
        wins = [(-1, 1, -1, 1), (0, 1, -1, 0), (0.25, 0.5, -0.5, -0.25), (0.05, 0.1, 0.05, 0.1), (0.1, 0.1, -0.2, -0.2),
                (-0.5, 0.5, 0.0, 0.0), (-4, 4, -4, 4), (-100, 100, -100, 100)]
        oscans, gscans = [], []
        for rot in rots:
            x, y = r * np.cos(a + rot + pose[2]), r * np.sin(a + rot + pose[2])
            oscans.append(ob.OracleScan(x, y, cartesian=True))
            gscans.append(sg.Scan(gpu, x, y, cartesian=True))
        sid, win, want = [], [], []
        for k, rot in enumerate(rots):
            for wn in wins:
                sid.append(k); win.append(wn)
                want.append(ob.orc.orc_match_bound(op, C.byref(oscans[k].s), C.byref(oparams), pose[0], pose[1], pose[2], rot, *wn))
        got = pyr.score_windows(gscans, sid, win, pose, gparams)
        assert np.array_equal(got, np.array(want))
        assert len(set(want)) > 10
        for s in gscans:
            s.close()
    finally:
        ob.orc.orc_pyramid_destroy(op)
        pyr.close(); gm.close()


@pytest.mark.parametrize("seed,lims", [(3300, (0.5, 0.5, 3.0, 0.5, 0.05)), (3301, (1.0, 0.6, 2.0, 1.0, 0.1)), (3302, (0.3, 0.3, 0.0, 0.5, 0.02))])
def test_m3rsm_matcher_matches_reference(sg, gpu, refso, seed, lims):
    """BruteForceMultiResolutionScanMatcher of the unmodified reference (libslamref.so) vs the host engine over K5:
    same pose delta, same probability"""
    rng = np.random.default_rng(seed)
    w = h = 256
    model = ob.CELL_MEAN
    rm = ob.RefMap(w, h, 0.05, model, ob.GROW_PLAIN, pyramid_oie=ob.OIE_DISCREPANCY)
    gm = sg.GridMap(gpu, w, h, 0.05, model, sg.GROW_PLAIN)
    pyr = sg.Pyramid(gpu, gm, sg.OIE_DISCREPANCY)
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    truth = np.array([0.3, -0.2, 0.1])
    try:
        for k in range(3):
            pose = truth + rng.normal(0, [0.1, 0.1, 0.05])
            r, a = room_scan(rng, 200, 2 * np.pi, pose=pose)
            occ = np.ones(200, np.uint8)
            refso.ref_append_scan(rm.h_, 200, ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.u8ptr(occ), pose[0], pose[1], pose[2], 1.0, 0,
                                  C.byref(oest), 0.3, np.inf, 0)
            gsc = sg.Scan(gpu, r, a)
            pyr.append_scan(gsc, pose, 1.0, 0, gest, blur=0.3)
            gsc.close()
        for lv in range(pyr.levels()):
            assert np.array_equal(pyr.level(lv), rm.export(lv)), lv
        xlim, ylim, rot_deg, ang_deg, tstep = lims
        r, a = room_scan(rng, 120, np.deg2rad(270), pose=truth, noise=0.003)
        init = truth + np.array([0.11, -0.07, np.deg2rad(0.8)])
        oparams = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY)
        m = ob.MatchResult()
        occ = np.ones(len(r), np.uint8)
        refso.ref_match_bf_m3rsm(rm.h_, len(r), ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.u8ptr(occ), ob.SPW_EVEN, C.byref(oparams),
                                 *init, xlim, ylim, np.deg2rad(rot_deg), np.deg2rad(ang_deg), tstep, C.byref(m))
        delta, prob, st = pyr.match_m3rsm(r, a, init, sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1), xlim, ylim,
                                          np.deg2rad(rot_deg), np.deg2rad(ang_deg), tstep)
        assert np.array_equal(delta, [m.dx, m.dy, m.dth]), (delta, (m.dx, m.dy, m.dth), st)
        assert prob == m.best_prob
        assert st["scored"] > 2 * st["rotations"] and st["branches"] > 3
    finally:
        pyr.close(); gm.close()
