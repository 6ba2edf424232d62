"""K2/K3 parity on the GPU: ray-cast cell lists bit-exact, map cells after scan insertion against the
CPU oracle.  Bar (north_star): cell probabilities within 1e-6 absolute; the kernels apply every cell's
updates sequentially in the reference's order, so the whole record array is asserted bit-equal."""
import ctypes as C

import numpy as np
import pytest

from helpers import room_scan
from oracle import binding as ob

pytestmark = pytest.mark.gpu
ATOL = 1e-6


def _point_quality(omqe, r, a):
    if omqe != ob.OMQE_AHR:
        return None
    v = np.zeros(len(r), np.uint32)
    ob.orc.orc_angle_histogram_values(len(r), ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), v.ctypes.data_as(C.POINTER(C.c_uint32)))
    return 1.0 / v


def test_raycast_lists_bit_exact(sg, gpu):
    rng = np.random.default_rng(2000)
    for scale, n in ((0.05, 181), (0.1, 360), (0.025, 97)):
        pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3))
        r, a = room_scan(rng, n, 2 * np.pi if n == 360 else np.deg2rad(270), pose=pose, noise=0.01)
        r[::17] = 0.0  # zero-length beams: a single cell
        om = ob.OracleMap(64, 64, scale, ob.CELL_LWW, ob.GROW_PLAIN)
        total, log = om.append_scan(ob.OracleScan(r, a), pose, log_cap=1 << 20)
        gm = sg.GridMap(gpu, 64, 64, scale, sg.CELL_LWW, sg.GROW_PLAIN)
        gsc = sg.Scan(gpu, r, a)
        offs, cells = gpu.raycast(gm, gsc, pose)
        assert offs[-1] == total == len(cells)
        assert np.array_equal(cells, log)
        assert (np.diff(offs) >= 1).all()
        gm.close(); gsc.close()


def test_raycast_segments_against_oracle(sg, gpu):
    rng = np.random.default_rng(2001)
    for sc in (0.05, 0.1, 1.0):
        segs = []
        for it in range(1500):
            b = rng.uniform(-5, 5, 2)
            e = b + rng.uniform(-6, 6, 2) * (0.02 if it % 7 == 0 else 1)
            if it % 11 == 0: e[0] = b[0]
            if it % 13 == 0: e[1] = b[1]
            if it % 17 == 0: b = np.round(b / sc) * sc
            if it % 19 == 0: e = np.round(e / sc) * sc
            if it % 23 == 0:
                b = (np.floor(b / sc) + 0.5) * sc
                e = (np.floor(e / sc) + 0.5) * sc
            segs.append([b[0], b[1], e[0], e[1]])
        offs, cells = gpu.raycast_segments(sc, segs)
        for i, s in enumerate(segs):
            want = ob.raycast(ob.orc.orc_raycast, *s, sc)
            assert np.array_equal(cells[offs[i]:offs[i + 1]], want), (sc, s)


def test_area_estimator_against_oracle(sg, gpu):
    rng = np.random.default_rng(2002)
    sc = 0.05
    oest = ob.estimator(ob.EST_AREA, shift=0.01 * 0.05)
    gest = sg.estimator(sg.EST_AREA, shift=0.01 * 0.05)
    beams, bounds, occ = [], [], []
    for it in range(20000):
        cx, cy = rng.integers(-5, 5, 2)
        cb, ct, cl, cr = sc * cy, sc * (cy + 1), sc * cx, sc * (cx + 1)
        b = rng.uniform(-0.4, 0.4, 2); e = rng.uniform(-0.4, 0.4, 2)
        m = it % 10
        if m == 0: e = np.array([rng.uniform(cl, cr), rng.uniform(cb, ct)])
        if m == 1: b = np.array([rng.uniform(cl, cr), rng.uniform(cb, ct)])
        if m == 2: e = np.array([cl, cb]) if it % 20 < 10 else np.array([cr, ct])
        if m == 3: b[1] = e[1] = cb if it % 20 < 10 else ct
        if m == 4: b[0] = e[0] = cl if it % 20 < 10 else cr
        if m == 5: e = np.array([cl, rng.uniform(cb, ct)])
        if m == 6:
            b = np.round(b / sc) * sc; e = np.round(e / sc) * sc
        if m == 7: e = np.array([rng.uniform(cl, cr), ct])
        beams.append([b[0], b[1], e[0], e[1]]); bounds.append([cb, ct, cl, cr]); occ.append(it % 2)
    got = gpu.estimate_occupancy(gest, beams, bounds, occ)
    want = np.zeros(2)
    for i in range(len(beams)):
        ob.orc.orc_estimate_occupancy(C.byref(oest), *beams[i], *bounds[i], occ[i], ob.dptr(want))
        assert np.array_equal(got[i], want, equal_nan=True), (i, beams[i], bounds[i], occ[i])


@pytest.mark.parametrize("model,est_type,blur,grow", [
    (ob.CELL_MEAN, ob.EST_CONST, 0.5, ob.GROW_NONE),
    (ob.CELL_TBM_CONSISTENT, ob.EST_AREA, 0.3, ob.GROW_PLAIN),
    (ob.CELL_TBM_UNKNOWN_EVEN, ob.EST_CONST, 0.0, ob.GROW_PLAIN),
    (ob.CELL_AFFINE, ob.EST_AREA, 0.2, ob.GROW_NONE),
    (ob.CELL_LWW, ob.EST_AREA, -0.01, ob.GROW_PLAIN),
    (ob.CELL_GMAPPING, ob.EST_CONST, 0.0, ob.GROW_TILED),
    (ob.CELL_CREDIBILIST, ob.EST_AREA, 0.3, ob.GROW_PLAIN),
])
def test_append_scan_matches_oracle(sg, gpu, model, est_type, blur, grow):
    rng = np.random.default_rng(2100 + model)
    w = h = 120 if grow != ob.GROW_NONE else 240
    om = ob.OracleMap(w, h, 0.05, model, grow)
    gm = sg.GridMap(gpu, w, h, 0.05, model, grow)
    tbm = model in (3, 4)
    kw = dict(occ=(0.95, 0.04) if tbm else (0.95, 1.0), empty=(0.01, 0.003) if tbm else (0.01, 1.0), shift=0.01 * 0.05)
    oest, gest = ob.estimator(est_type, **kw), sg.estimator(est_type, **kw)
    for k in range(3):
        pose = (rng.uniform(-1, 1), rng.uniform(-1, 1), rng.uniform(-3, 3))
        r, a = room_scan(rng, 241, np.deg2rad(270), pose=pose, noise=0.01)
        occ = (rng.random(241) < 0.95).astype(np.uint8)
        omqe = ob.OMQE_AHR if k == 2 else ob.OMQE_IDLE
        margin = 2 if k == 1 else 0
        n1, _ = om.append_scan(ob.OracleScan(r, a, occ=occ), pose, 0.9, margin, oest, blur=blur, max_range=5.5, omqe=omqe)
        gsc = sg.Scan(gpu, r, a, occ=occ)
        n2 = gpu.append_scan(gm, gsc, pose, 0.9, margin, gest, blur=blur, max_range=5.5, point_quality=_point_quality(omqe, r, a))
        assert n1 == n2
        assert gm.info() == om.info()
        got, want = gm.download(), om.cells()
        assert np.allclose(got[..., 0], want[..., 0], rtol=0, atol=ATOL, equal_nan=True)  # the stated bar
        assert np.array_equal(got, want, equal_nan=True)                                   # what the kernels deliver
        lut, unk = gm.lut(0)
        olut, ounk = om.lut(0)
        if model != ob.CELL_GMAPPING:
            assert np.array_equal(lut, olut, equal_nan=True) and unk == ounk
        gsc.close()
    assert (om.cells()[..., 0] != om.cells()[0, 0, 0]).sum() > 1000
    gm.close()


def test_append_scan_bounded_map_drops_outside_updates(sg, gpu):
    rng = np.random.default_rng(2200)
    om = ob.OracleMap(60, 60, 0.1, ob.CELL_MEAN, ob.GROW_NONE)   # 6 m map, 8 x 6 m room: beams leave the map
    gm = sg.GridMap(gpu, 60, 60, 0.1, sg.CELL_MEAN, sg.GROW_NONE)
    r, a = room_scan(rng, 360, 2 * np.pi, pose=(0.3, 0.2, 0.1))
    n1, _ = om.append_scan(ob.OracleScan(r, a), (0.3, 0.2, 0.1), 1.0, 0, ob.estimator(), blur=0.3)
    gsc = sg.Scan(gpu, r, a)
    n2 = gpu.append_scan(gm, gsc, (0.3, 0.2, 0.1), 1.0, 0, sg.estimator(), blur=0.3)
    assert n1 == n2 and gm.info() == om.info()
    assert np.array_equal(gm.download(), om.cells())
    # empty scan and a scan shorter than twice the margin are no-ops
    empty = sg.Scan(gpu, np.zeros(0), np.zeros(0))
    assert gpu.append_scan(gm, empty, (0, 0, 0)) == 0
    short = sg.Scan(gpu, np.ones(3), np.zeros(3))
    assert gpu.append_scan(gm, short, (0, 0, 0), margin=2) == 0
    assert np.array_equal(gm.download(), om.cells())
    gm.close(); gsc.close(); empty.close(); short.close()


@pytest.mark.parametrize("model", [ob.CELL_LWW, ob.CELL_AFFINE, ob.CELL_MEAN, ob.CELL_TBM_CONSISTENT, ob.CELL_TBM_UNKNOWN_EVEN,
                                   ob.CELL_GMAPPING, ob.CELL_CREDIBILIST])
def test_single_cell_update_reset_read(sg, gpu, model):
    rng = np.random.default_rng(2300 + model)
    for grow in (ob.GROW_PLAIN, ob.GROW_TILED):
        om = ob.OracleMap(20, 14, 0.1, model, grow)
        gm = sg.GridMap(gpu, 20, 14, 0.1, model, grow)
        for k in range(60):
            x, y = int(rng.integers(-25, 25)), int(rng.integers(-20, 20))
            p, q, quality = rng.random(3)
            if k % 9 == 0: p = np.nan
            obst = rng.uniform(-2, 2, 2)
            ob.orc.orc_map_ensure_inside(om.h_, x, y)
            i = om.info()
            rec = om.cells()[y + i["oy"], x + i["ox"]]
            ob.orc.orc_cell_update(model, ob.dptr(rec), 1, p, q, obst[0], obst[1], quality)
            gm.update_cell(x, y, True, p, q, obst, quality)
            assert gm.info() == om.info()
            assert np.array_equal(gm.read_cell(x, y), rec, equal_nan=True)
        assert np.array_equal(gm.download(), om.cells(), equal_nan=True)
        proto = gm.read_cell(10 ** 6, 10 ** 6)  # far outside: the unknown prototype
        unk = np.zeros(8)
        ob.orc.orc_default_unknown(model, ob.dptr(unk))
        assert np.array_equal(proto, unk[:len(proto)])
        gm.reset_cell(3, -2, proto)
        assert np.array_equal(gm.read_cell(3, -2), proto)
        gm.close()


def test_viny_shape_full_size_update(sg, gpu):
    """BASELINE config 2 shape: 800x800 @0.05, TBM cells, area estimator, blur 0.3, 1081 beams over 270 deg"""
    rng = np.random.default_rng(2400)
    kw = dict(occ=(0.95, 0.04), empty=(0.01, 0.003), shift=0.01 * 0.05)
    oest, gest = ob.estimator(ob.EST_AREA, **kw), sg.estimator(sg.EST_AREA, **kw)
    om = ob.OracleMap(800, 800, 0.05, ob.CELL_TBM_CONSISTENT, ob.GROW_PLAIN)
    gm = sg.GridMap(gpu, 800, 800, 0.05, sg.CELL_TBM_CONSISTENT, sg.GROW_PLAIN)
    total = 0
    for k in range(4):
        pose = (rng.uniform(-3, 3), rng.uniform(-3, 3), rng.uniform(-3, 3))
        r, a = room_scan(rng, 1081, np.deg2rad(270), half_w=15.0, half_h=12.0, pose=pose, noise=0.01)
        n1, _ = om.append_scan(ob.OracleScan(r, a), pose, 0.9, 0, oest, blur=0.3)
        gsc = sg.Scan(gpu, r, a)
        n2 = gpu.append_scan(gm, gsc, pose, 0.9, 0, gest, blur=0.3)
        assert n1 == n2
        total += n2
        gsc.close()
    assert total > 4e5
    assert gm.info() == om.info()
    assert np.array_equal(gm.download(), om.cells(), equal_nan=True)
    gm.close()


def test_filter_scan_matches_oracle(sg, gpu):
    rng = np.random.default_rng(2500)
    for grow in (ob.GROW_NONE, ob.GROW_PLAIN):
        om = ob.OracleMap(60, 60, 0.1, ob.CELL_LWW, grow)
        gm = sg.GridMap(gpu, 60, 60, 0.1, sg.CELL_LWW, grow)
        r = rng.uniform(0.2, 6, 200); a = np.linspace(-2, 2, 200)
        occ = (rng.random(200) < 0.8).astype(np.uint8)
        for skip, mr in ((0, -1.0), (3, -1.0), (0, 3.0), (2, 2.5)):
            k1 = np.zeros(200, np.int32)
            n1 = ob.orc.orc_filter_scan(om.h_, 200, ob.dptr(r), ob.dptr(a), ob.u8ptr(occ), 0.3, 0.2, 0.5, skip, mr, ob.iptr(k1))
            got = gm.filter_scan(r, a, (0.3, 0.2, 0.5), occ=occ, skip_rate=skip, max_range=mr)
            assert np.array_equal(got, k1[:n1])
        gm.close()


def test_library_division_is_correctly_rounded(sg, gpu):
    """the TBM update chain divides through a shortcut for subnormal numerators (sg::div_chain): against IEEE division on the
    host, over random operands, the subnormal range, ties on the subnormal grid and denominators within rounding of one"""
    rng = np.random.default_rng(77)
    n = 400000
    a = np.ldexp(rng.random(n) + 0.5, rng.integers(-1080, -880, n))          # tiny and subnormal numerators
    b = np.ldexp(rng.random(n) + 0.5, rng.integers(-4, 4, n))
    a[:1000] = 5e-324 * rng.integers(1, 40, 1000)                             # the bottom of the subnormal range
    b[:500] = 1.0
    b[500:1000] = 1.0 - 2.0 ** -53 * rng.integers(1, 8, 500)
    a[1800:4000] = 5e-324 * rng.integers(1, 2 ** 20, 2200)                     # k units against b near one: around the shortcut's edge
    b[1800:4000] = 1.0 + rng.normal(0, 1, 2200) * 2.0 ** -rng.integers(2, 30, 2200)
    # exact ties: numerator = odd multiple of half the smallest subnormal, scaled by a power-of-two denominator
    a[1000:1400] = 5e-324 * (2 * rng.integers(1, 1000, 400) + 1)
    b[1000:1400] = 2.0
    a[1400:1800] = np.ldexp(2 * rng.integers(1, 2 ** 40, 400) + 1.0, -1074 - 40 + 10)
    b[1400:1800] = 2.0 ** 11
    sign = rng.integers(0, 2, n) * 2 - 1
    a2 = np.concatenate([a * sign, rng.normal(0, 1, 50000), np.ldexp(rng.random(50000), rng.integers(-1000, 1000, 50000))])
    b2 = np.concatenate([b, rng.normal(0, 1, 50000), np.ldexp(rng.random(50000) + 0.5, rng.integers(-1000, 1000, 50000))])
    with np.errstate(all="ignore"):
        want = a2 / b2
    got = gpu.debug_div(a2, b2)
    bad = ~((got == want) | (np.isnan(got) & np.isnan(want)))
    assert not bad.any(), (a2[bad][:5], b2[bad][:5], got[bad][:5], want[bad][:5])
    assert (np.abs(want[:n]) < 2.3e-308).sum() > 100000  # the subnormal results really were exercised


@pytest.mark.parametrize("model", [ob.CELL_TBM_CONSISTENT, ob.CELL_TBM_UNKNOWN_EVEN, ob.CELL_CREDIBILIST])
@pytest.mark.parametrize("est_type", [ob.EST_CONST, ob.EST_AREA])
def test_tbm_cells_through_the_subnormal_regime(sg, gpu, model, est_type):
    """a robot that stands still: the cells around it take thousands of "empty" observations, their occupied / unknown
    masses decay through the subnormal range and stick at the smallest subnormals.  Every record stays bit-equal to the
    oracle's through the transient and in the saturated state (division shortcut, skipped repeats, robot-cell side chain)"""
    rng = np.random.default_rng(2100 + model + 10 * est_type)
    pose = np.array([0.317, -0.223, 0.1])
    r, a = room_scan(rng, 721, 2 * np.pi, half_w=3.0, half_h=2.5, pose=pose, noise=0.0)
    om = ob.OracleMap(160, 160, 0.05, model, ob.GROW_PLAIN)
    gm = sg.GridMap(gpu, 160, 160, 0.05, model, sg.GROW_PLAIN)
    oe = ob.estimator(est_type, occ=(0.95, 0.9), empty=(0.01, 0.8), low_qual=0.01, unknown_qual=0.5, shift=0.0005)
    ge = sg.estimator(est_type, occ=(0.95, 0.9), empty=(0.01, 0.8), low_qual=0.01, unknown_qual=0.5, shift=0.0005)
    gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
    saw_subnormal = False
    for k in range(7):
        p = pose + (np.array([0.004, -0.003, 0.01]) if k == 4 else 0.0)  # one scan from a slightly different pose
        want, _ = om.append_scan(osc, p, 1.0, 0, oe, blur=0.2)
        got = gpu.append_scan(gm, gsc, p, 1.0, 0, ge, blur=0.2)
        assert got == want
        cells_o, cells_g = om.cells(), gm.download()
        assert np.array_equal(cells_g, cells_o, equal_nan=True), k
        masses = np.abs(cells_o[..., 2:5])
        saw_subnormal |= bool(((masses > 0) & (masses < 2.3e-308)).any())
    if est_type == ob.EST_CONST:  # (the area estimator's qualities decay the masses too slowly for seven scans)
        assert saw_subnormal, "the scenario never reached the subnormal range"
    gm.close(); gsc.close()
