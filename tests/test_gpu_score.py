"""K1 parity on the GPU: libslamgpu.so (through the C ABI) against the CPU oracle.

Bars (BASELINE.json north_star): selected candidate index bit-exact; per-pose scores within
1e-5 relative.  The kernels are written to be bit-exact for every mode built only from
+ - * / (obstacle, max, mean, overlap), so those are asserted with array_equal; the GMapping
mode goes through exp() and is asserted at rtol 1e-5 (RTOL below).
"""
import ctypes as C

import numpy as np
import pytest

from helpers import random_cells, room_map_cells, room_scan
from oracle import binding as ob

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _setup(sg, gpu, rng, model, n_pts, size=200, scale=0.05, passes=3, spw=ob.SPW_EVEN, factor=False, fov=270):
    cells = room_map_cells(rng, size, size, scale, model, passes=passes)
    om = ob.OracleMap(size, size, scale, model, ob.GROW_NONE)
    om.set_cells(cells)
    gm = sg.GridMap(gpu, size, size, scale, model)
    gm.upload(cells)
    pose0 = np.array([0.2, -0.1, 0.3])
    r, a = room_scan(rng, n_pts, np.deg2rad(fov), pose=pose0, noise=0.01)
    w = np.empty(n_pts)
    ob.orc.orc_point_weights(spw, n_pts, ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.dptr(w))
    f = rng.uniform(0.5, 1.5, n_pts) if factor else None
    osc = ob.OracleScan(r, a, weight=w, factor=f)
    gsc = sg.Scan(gpu, r, a, weight=w, factor=f)
    return om, gm, osc, gsc, pose0


def _argbest(scores, init):
    best = C.c_double()
    idx = ob.orc.orc_argbest(ob.dptr(ob.f64(scores)), len(scores), init, C.byref(best))
    return idx, best.value


@pytest.mark.parametrize("model,oie", [(ob.CELL_LWW, 0), (ob.CELL_MEAN, 0), (ob.CELL_TBM_CONSISTENT, 0),
                                       (ob.CELL_TBM_UNKNOWN_EVEN, 1), (ob.CELL_AFFINE, 1), (ob.CELL_CREDIBILIST, 0)])
@pytest.mark.parametrize("trig", [0, 1])
def test_list_obstacle_bit_exact(sg, gpu, model, oie, trig):
    rng = np.random.default_rng(1000 + model * 3 + trig)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, model, 181, spw=[ob.SPW_EVEN, ob.SPW_VINY, ob.SPW_AHR][model % 3],
                                  factor=(model == ob.CELL_MEAN))
    poses = p0 + rng.normal(0, [0.2, 0.2, 0.1], (700, 3))
    poses[::50] += [30.0, -30.0, 0.0]  # far outside the map: unknown cells
    want = om.score(osc, ob.spe_params(ob.OOPE_OBSTACLE, oie), poses)
    init = float(np.sort(want)[len(want) // 2])
    got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(sg.OOPE_OBSTACLE, oie, trig=trig), poses, init_score=init)
    lut, unk = gm.lut(oie)
    olut, ounk = om.lut(oie)
    assert np.array_equal(lut, olut) and unk == ounk
    assert np.array_equal(got, want)
    assert (idx, best) == _argbest(want, init)
    gm.close(); gsc.close()


@pytest.mark.parametrize("oope", [ob.OOPE_MAX, ob.OOPE_MEAN, ob.OOPE_OVERLAP])
def test_list_window_modes_bit_exact(sg, gpu, oope):
    rng = np.random.default_rng(1100 + oope)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, ob.CELL_MEAN, 90, fov=240)
    poses = p0 + rng.normal(0, [0.2, 0.2, 0.1], (130, 3))
    for win in ((0.1, 0.1), (0.05, 0.2), (0.0, 0.0), (0.3, 0.0)):
        want = om.score(osc, ob.spe_params(oope, 0, win_v=win[0], win_h=win[1]), poses)
        for trig in (0, 1):
            # the overlap OOPE depends continuously on the point position, so the library always takes
            # libm (host) trig for it, whatever trig_mode asks: still bit-exact
            got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(oope, 0, win_v=win[0], win_h=win[1], trig=trig), poses)
            assert np.array_equal(got, want), (win, trig)
            assert (idx, best) == _argbest(want, -np.inf)
    gm.close(); gsc.close()


def test_list_prerotated_cartesian(sg, gpu):
    rng = np.random.default_rng(1200)
    om, gm, _, _, p0 = _setup(sg, gpu, rng, ob.CELL_MEAN, 50)
    r, a = room_scan(rng, 150, np.deg2rad(270), pose=p0)
    x, y = r * np.cos(a + p0[2]), r * np.sin(a + p0[2])
    osc = ob.OracleScan(x, y, cartesian=True)
    gsc = sg.Scan(gpu, x, y, cartesian=True)
    poses = p0 + rng.normal(0, [0.3, 0.3, 0.0], (257, 3))
    for oope, win in ((ob.OOPE_OBSTACLE, (0, 0)), (ob.OOPE_MAX, (0.2, 0.1))):
        want = om.score(osc, ob.spe_params(oope, 0, win_v=win[0], win_h=win[1], prerotated=1), poses)
        got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(oope, 0, win_v=win[0], win_h=win[1], prerotated=1), poses)
        assert np.array_equal(got, want)
        assert (idx, best) == _argbest(want, -np.inf)
    # a Cartesian scan scored NOT pre-rotated goes through range/angle (sensor_data.h:47-70)
    want = om.score(osc, ob.spe_params(ob.OOPE_OBSTACLE, 0), poses)
    got, _, _ = gpu.score_poses(gm, gsc, sg.spe_params(sg.OOPE_OBSTACLE, 0, trig=sg.TRIG_HOST), poses)
    assert np.array_equal(got, want)
    gm.close(); gsc.close()


def test_list_gmapping_oope(sg, gpu):
    rng = np.random.default_rng(1300)
    cells = room_map_cells(rng, 200, 200, 0.05, ob.CELL_GMAPPING, passes=4)
    om = ob.OracleMap(200, 200, 0.05, ob.CELL_GMAPPING)
    om.set_cells(cells)
    gm = sg.GridMap(gpu, 200, 200, 0.05, sg.CELL_GMAPPING)
    gm.upload(cells)
    r, a = room_scan(rng, 360, 2 * np.pi, pose=(0.1, 0.1, 0.0), noise=0.005)
    osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
    poses = np.array([0.1, 0.1, 0.0]) + rng.normal(0, [0.05, 0.05, 0.02], (300, 3))
    want = om.score(osc, ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1), poses)
    got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1), poses)
    assert (want > 0).sum() > 250
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    # the reference's 1-entry cell cache (quirk Q7), restarted per pose
    want_c = np.array([om.score(osc, ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1), p[None],
                                cache=ob.GmCache(0, 0, -1.0))[0] for p in poses[:60]])
    got_c, _, _ = gpu.score_poses(gm, gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=1), poses[:60])
    np.testing.assert_allclose(got_c, want_c, rtol=RTOL, atol=0)
    # the cache as the estimator object really keeps it: alive across poses (and calls)
    params_o = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    params_g = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2)
    cache = ob.GmCache(0, 0, -1.0)
    # a scan that ends where it starts + tiny pose steps: a candidate's first point mostly falls into the cell of the
    # previous candidate's last point and takes its (stale) probability
    gsc.close()
    r, a = np.append(r, r[0]), np.append(a, a[0])
    osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
    seq = np.array([0.1, 0.1, 0.0]) + np.cumsum(rng.normal(0, [0.004, 0.004, 0.001], (120, 3)), axis=0)
    want_1 = om.score(osc, params_o, seq[:70], cache=cache)
    mid = ob.GmCache(cache.cx, cache.cy, cache.prob)
    want_2 = om.score(osc, params_o, seq[70:], cache=cache)
    fresh = np.array([om.score(osc, params_o, q[None], cache=ob.GmCache(0, 0, -1.0))[0] for q in seq])
    assert (np.concatenate([want_1, want_2]) != fresh).sum() > 10  # the carried cache really changes scores
    got_1, st_1 = gpu.score_poses_chained(gm, gsc, params_g, seq[:70])
    np.testing.assert_allclose(got_1, want_1, rtol=1e-12, atol=0)
    assert (st_1[-1].cx, st_1[-1].cy) == (mid.cx, mid.cy) and abs(st_1[-1].prob - mid.prob) <= 1e-12 * abs(mid.prob)
    got_2, st_2 = gpu.score_poses_chained(gm, gsc, params_g, seq[70:], state=st_1[-1])
    np.testing.assert_allclose(got_2, want_2, rtol=1e-12, atol=0)
    assert (st_2[-1].cx, st_2[-1].cy) == (cache.cx, cache.cy)
    # a consumer that stops early continues from the state after the pose it stopped at
    got_3, _ = gpu.score_poses_chained(gm, gsc, params_g, seq[30:70], state=st_1[29])
    np.testing.assert_allclose(got_3, want_1[30:], rtol=1e-12, atol=0)
    with pytest.raises(sg.SlamGpuError):
        gpu.score_poses(gm, gsc, params_g, seq[:5])  # the plain entry point has no cache to carry
    gm.close(); gsc.close()


def test_chained_cache_through_single_cell_poses(sg, gpu):
    """degenerate chains: scans that fall into ONE cell keep the cache of an earlier pose alive across whole poses"""
    rng = np.random.default_rng(1301)
    cells = room_map_cells(rng, 120, 120, 0.05, ob.CELL_GMAPPING, passes=3)
    om = ob.OracleMap(120, 120, 0.05, ob.CELL_GMAPPING); om.set_cells(cells)
    gm = sg.GridMap(gpu, 120, 120, 0.05, sg.CELL_GMAPPING); gm.upload(cells)
    # 5 beams inside half a degree at ~1 m: all end in the same cell for most poses
    r = np.full(5, 1.0) + rng.normal(0, 1e-4, 5)
    a = np.linspace(-0.004, 0.004, 5)
    osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
    seq = np.array([0.2, 0.1, 0.3]) + np.cumsum(rng.normal(0, [0.003, 0.003, 0.0005], (200, 3)), axis=0)
    po = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.0, gm_window=1)
    pg = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.0, gm_window=1, gm_cache=2)
    cache = ob.GmCache(0, 0, -1.0)
    want = om.score(osc, po, seq, cache=cache)
    got, st = gpu.score_poses_chained(gm, gsc, pg, seq)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-300)
    assert (st[-1].cx, st[-1].cy) == (cache.cx, cache.cy)
    gm.close(); gsc.close()


def _bf_axes(base, args):
    xs, ys, ts = np.empty(4096), np.empty(4096), np.empty(4096)
    nx, ny, nt = C.c_int32(), C.c_int32(), C.c_int32()
    n = ob.orc.orc_bf_enumerate(*base, *args, None, 0, ob.dptr(xs), nx, ob.dptr(ys), ny, ob.dptr(ts), nt)
    P = np.empty((n, 3))
    ob.orc.orc_bf_enumerate(*base, *args, ob.dptr(P), n, None, None, None, None, None, None)
    return xs[:nx.value].copy(), ys[:ny.value].copy(), ts[:nt.value].copy(), P


@pytest.mark.parametrize("trig", [0, 1])
@pytest.mark.parametrize("model,factor", [(ob.CELL_MEAN, False), (ob.CELL_TBM_CONSISTENT, True)])
def test_grid_matches_oracle_and_list(sg, gpu, trig, model, factor):
    rng = np.random.default_rng(1400 + trig + model)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, model, 121, factor=factor, spw=ob.SPW_VINY)
    base = p0 + np.array([0.07, -0.04, 0.02])
    xs, ys, ts, P = _bf_axes(base, (-0.5, 0.5, 0.05, -0.3, 0.3, 0.05, -0.1, 0.1, 0.02))
    assert len(P) == len(xs) * len(ys) * len(ts) and len(ys) % 8 != 0
    want = om.score(osc, ob.spe_params(), P)
    init = om.score(osc, ob.spe_params(), base[None])[0]
    got, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(trig=trig), xs, ys, ts, init_score=init)
    assert np.array_equal(got, want)
    assert (idx, best) == _argbest(want, init)
    got2, idx2, best2 = gpu.score_poses(gm, gsc, sg.spe_params(trig=trig), P, init_score=init)
    assert np.array_equal(got2, want) and (idx2, best2) == (idx, best)
    st = gpu.score_stats()
    assert st["variant"] in (0, 3, 4) and st["evals"] == len(P) * 121  # list kernel, or its small-batch forms
    gm.close(); gsc.close()


def test_argmax_ties_and_accept_rule(sg, gpu):
    """constant map: every candidate ties -> the lowest index wins, and only if it beats init_score"""
    cells = np.zeros((50, 50, 2))
    cells[..., 0] = 0.7; cells[..., 1] = 3
    gm = sg.GridMap(gpu, 50, 50, 0.1, sg.CELL_MEAN)
    gm.upload(cells)
    r = np.full(40, 1.0); a = np.linspace(-1, 1, 40)
    gsc = sg.Scan(gpu, r, a)
    xs = np.linspace(-0.5, 0.5, 33); ys = np.linspace(-0.5, 0.5, 19); ts = np.linspace(-0.2, 0.2, 5)
    got, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts, init_score=0.1)
    assert len(set(got.tolist())) == 1 and idx == 0 and best == got[0]
    _, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts, init_score=got[0])  # ties never accepted
    assert idx == -1 and best == got[0]
    # one better cell row makes a unique maximum somewhere in the middle; duplicates of it tie -> first wins
    P = np.zeros((1000, 3)); P[:, 0] = np.linspace(-0.4, 0.4, 1000)
    P[600:] = P[:400]
    got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(), P, init_score=-1.0)
    assert idx == 0
    gm.close(); gsc.close()


def test_empty_and_degenerate_inputs(sg, gpu):
    gm = sg.GridMap(gpu, 20, 20, 0.1, sg.CELL_LWW)
    gsc = sg.Scan(gpu, np.array([1.0, 2.0]), np.array([0.0, 0.5]))
    got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(), np.zeros((0, 3)), init_score=0.25)
    assert len(got) == 0 and idx == -1 and best == 0.25
    # an empty scan has zero total weight: the reference returns NaN (unknown probability), never accepted
    empty = sg.Scan(gpu, np.zeros(0), np.zeros(0))
    got, idx, best = gpu.score_poses(gm, empty, sg.spe_params(), np.zeros((5, 3)), init_score=0.25)
    assert np.isnan(got).all() and idx == -1 and best == 0.25
    # all-unknown map: every point reads the prototype
    got, idx, _ = gpu.score_poses(gm, gsc, sg.spe_params(), np.array([[0.0, 0.0, 0.0], [100.0, 5.0, 1.0]]))
    assert got[0] == got[1] == 0.5 and idx == 0
    with pytest.raises(sg.SlamGpuError):  # a pre-rotated scan carries its theta: no candidate grid over it
        gpu.score_grid(gm, gsc, sg.spe_params(prerotated=1), [0.0], [0.0], [0.0])
    gm.close(); gsc.close(); empty.close()


def test_guard_forces_host_trig_on_cell_borders(sg, gpu):
    """scan points that land exactly on cell borders must trip the device-trig guard and still be exact"""
    rng = np.random.default_rng(1500)
    cells = random_cells(rng, 80, 80, ob.CELL_MEAN, known_frac=1.0)
    om = ob.OracleMap(80, 80, 0.1, ob.CELL_MEAN); om.set_cells(cells)
    gm = sg.GridMap(gpu, 80, 80, 0.1, sg.CELL_MEAN); gm.upload(cells)
    r = np.array([1.0, 2.0, 0.5, 1.5]); a = np.array([0.0, np.pi / 2, np.pi, 0.0])  # axis-aligned beams
    osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
    poses = np.array([[0.1 * k, 0.1 * (k % 7), 0.0] for k in range(-20, 20)])  # multiples of the cell size
    want = om.score(osc, ob.spe_params(), poses)
    got, idx, best = gpu.score_poses(gm, gsc, sg.spe_params(trig=sg.TRIG_DEVICE), poses)
    assert gpu.score_stats()["guard_hits"] > 0
    assert np.array_equal(got, want) and (idx, best) == _argbest(want, -np.inf)
    gm.close(); gsc.close()


def test_full_size_bruteforce_properties(sg, gpu):
    """BASELINE config 3 shape (2000x2000 grid, 1081 beams, 101x101x100 candidates): sampled scores against
    the oracle, the arg-max against numpy over all GPU scores, and list-vs-grid agreement on a slice."""
    rng = np.random.default_rng(1600)
    size, scale, n = 2000, 0.05, 1081
    cells = room_map_cells(rng, size, size, scale, ob.CELL_MEAN, half_w=30.0, half_h=22.0, passes=2)
    om = ob.OracleMap(size, size, scale, ob.CELL_MEAN); om.set_cells(cells)
    gm = sg.GridMap(gpu, size, size, scale, sg.CELL_MEAN); gm.upload(cells)
    p0 = np.array([1.3, -2.1, 0.4])
    r, a = room_scan(rng, n, np.deg2rad(270), half_w=30.0, half_h=22.0, pose=p0, noise=0.01)
    osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
    xs, ys, ts, P = _bf_axes(p0 + [0.04, -0.03, 0.01], (-1, 1, 0.02, -1, 1, 0.02, -0.5, 0.5, 0.01))
    assert (len(xs), len(ys), len(ts)) == (101, 101, 100)
    got, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
    assert not np.isnan(got).any()
    assert idx == int(np.argmax(got)) and best == got[idx]  # np.argmax returns the first maximum
    pick = rng.choice(len(P), 3000, replace=False)
    want = om.score(osc, ob.spe_params(), P[pick])
    assert np.array_equal(got[pick], want)
    sl = slice(400000, 420000)
    got_l, _, _ = gpu.score_poses(gm, gsc, sg.spe_params(), P[sl])
    assert np.array_equal(got_l, got[sl])
    gm.close(); gsc.close()


@pytest.mark.parametrize("max_variant", [5, 4, 3, 2])
@pytest.mark.parametrize("nx,ny,nt,ystep,expect_R,fits_box,fits_nibbles", [
    (33, 19, 5, 0.02, 2, True, True),         # few candidates: 2 rows per thread (v4: point factors -> v2)
    (400, 160, 8, 0.02, 4, True, True),       # medium: 4 rows per thread (v4: always 8)
    (300, 400, 12, 0.013, 8, True, True),     # many: 8 rows per thread
    (101, 53, 7, 0.049, 2, True, True),       # y step just under a cell: 8 distinct rows per thread (v4 with DMAX = 8)
    (37, 21, 3, 0.031, 2, True, True),        # leftover columns only from the second band on; DMAX = 6
    (21, 10, 3, 0.6, 8, False, False),        # y values far apart: no TMA box, no nibble deltas -> explicit row table (v1)
    (40, 17, 4, -0.03, 2, True, True),        # decreasing y: negative row deltas (v4 -> v2)
])
def test_grid_kernel_variants(sg, gpu, max_variant, nx, ny, nt, ystep, expect_R, fits_box, fits_nibbles):
    if max_variant == 3 and nx in (101, 37):
        pytest.skip("the two v4-specific shapes are not run through the experimental TMA variant")
    rng = np.random.default_rng(1700 + nx)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, ob.CELL_MEAN, 73, spw=ob.SPW_AHR, factor=(nx == 33))
    xs = p0[0] + 0.011 * (np.arange(nx) - nx // 2)
    ys = p0[1] + ystep * (np.arange(ny) - ny // 2)
    ts = p0[2] + 0.017 * (np.arange(nt) - nt // 2)
    P = np.stack(np.meshgrid(ts, ys, xs, indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
    gpu.set_option("grid_variant", max_variant)
    try:
        got, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
    finally:
        gpu.set_option("grid_variant", 0)
    st = gpu.score_stats()
    pick = rng.choice(len(P), min(len(P), 4000), replace=False)
    want = om.score(osc, ob.spe_params(), P[pick])
    assert np.array_equal(got[pick], want)
    assert idx == int(np.argmax(got)) and best == got[idx]
    v4 = max_variant >= 4 and ystep > 0 and fits_nibbles and nx != 33
    v5 = v4 and max_variant == 5 and int(np.floor(7 * ystep / 0.05)) + 2 <= 4  # the 4-row patch of the cp.async pipeline
    expect_variant = 5 if v5 else 4 if v4 else (1 if not fits_nibbles else (3 if (max_variant == 3 and fits_box) else 2))
    assert st["variant"] == expect_variant and st["rows_per_thread"] == (8 if v4 else expect_R), st
    gm.close(); gsc.close()


def test_grid_row_dedupe_kernels_against_the_general_one(sg, gpu):
    """k_score_grid5 (default: cp.async pipeline) and k_score_grid4 against k_score_grid2 on the same candidate grids, every
    score: all DMAX instantiations (y steps from a tenth of a cell to a full cell), column counts around the band widths,
    uneven point weights (the weight pairs ride in the copy), odd and tiny beam counts"""
    rng = np.random.default_rng(1750)
    for nx, ny, nt, ystep, spw, n_pts in [(32, 8, 2, 0.004, ob.SPW_EVEN, 97), (65, 23, 3, 0.012, ob.SPW_VINY, 96),
                                          (31, 40, 2, 0.02, ob.SPW_EVEN, 1), (101, 101, 2, 0.02, ob.SPW_EVEN, 97),
                                          (96, 9, 4, 0.027, ob.SPW_AHR, 2), (70, 33, 2, 0.035, ob.SPW_EVEN, 3),
                                          (45, 17, 3, 0.0499, ob.SPW_VINY, 97), (5, 64, 3, 0.02, ob.SPW_EVEN, 4),
                                          (61, 30, 2, 0.02, ob.SPW_VINY, 5)]:
        om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, ob.CELL_TBM_CONSISTENT, n_pts, spw=spw)
        xs = p0[0] + 0.013 * (np.arange(nx) - nx // 2)
        ys = p0[1] + ystep * (np.arange(ny) - ny // 2)
        ts = p0[2] + 0.02 * (np.arange(nt) - nt // 2)
        dmax = int(np.floor(7 * ystep / 0.05)) + 2
        got5, idx5, best5 = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
        assert gpu.score_stats()["variant"] == (5 if dmax <= 4 else 4)
        res = {}
        for v in (4, 2):
            gpu.set_option("grid_variant", v)
            try:
                res[v] = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
            finally:
                gpu.set_option("grid_variant", 0)
            assert gpu.score_stats()["variant"] == v
        for v in (4, 2):
            assert np.array_equal(got5, res[v][0]) and (idx5, best5) == res[v][1:]
        P = np.stack(np.meshgrid(ts, ys, xs, indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
        pick = rng.choice(len(P), min(len(P), 1500), replace=False)
        assert np.array_equal(got5[pick], om.score(osc, ob.spe_params(), P[pick]))
        gm.close(); gsc.close()


@pytest.mark.parametrize("trig", [0, 1])
@pytest.mark.parametrize("mode,win", [("max", (0.1, 0.1)), ("max", (0.15, 0.1)), ("mean", (0.1, 0.1)), ("mean", (0.12, 0.07)),
                                      ("mean", (0.26, 0.31)), ("max", (0.0, 0.1)), ("overlap", (0.1, 0.1)), ("gmapping", None)])
def test_candidate_grid_with_every_oope(sg, gpu, mode, win, trig):
    """slamgpu_stage_grid with the window OOPEs (occupancy_observation_probability.h:29-99) and the GMapping OOPE: max / mean
    are scored by the grid kernel out of the map's window LUTs (one gather per evaluation), overlap / GMapping / degenerate
    windows as a device-expanded pose list; every score against the oracle and against slamgpu_score_poses"""
    rng = np.random.default_rng(1760 + trig)
    model = ob.CELL_GMAPPING if mode == "gmapping" else ob.CELL_TBM_CONSISTENT
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, model, 83, spw=ob.SPW_VINY)
    if mode == "gmapping":
        po, pg = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1), sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, trig=trig)
    else:
        code = dict(max=ob.OOPE_MAX, mean=ob.OOPE_MEAN, overlap=ob.OOPE_OVERLAP)[mode]
        po, pg = ob.spe_params(code, win_v=win[0], win_h=win[1]), sg.spe_params(code, win_v=win[0], win_h=win[1], trig=trig)
    xs = p0[0] + 0.02 * (np.arange(41) - 20); ys = p0[1] + 0.02 * (np.arange(29) - 14); ts = p0[2] + 0.01 * (np.arange(5) - 2)
    # a few candidates far outside the map: windows over unknown cells on every side
    xs[0] -= 9.0; xs[-1] += 9.0; ys[0] -= 9.0; ys[-1] += 9.0
    P = np.stack(np.meshgrid(ts, ys, xs, indexing="ij"), -1).reshape(-1, 3)[:, ::-1]
    want = om.score(osc, po, P)
    got, idx, best = gpu.score_grid(gm, gsc, pg, xs, ys, ts)
    st = gpu.score_stats()
    tabulated = mode in ("max", "mean") and win[0] > 0
    assert (st["variant"] == 1) if tabulated else (st["variant"] in (0, 3, 4)), st   # grid kernel on the window tables / list kernels
    if mode == "gmapping":
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    else:
        assert np.array_equal(got, want)
        assert idx == int(np.argmax(want)) and best == want[idx]
    got2, idx2, best2 = gpu.score_poses(gm, gsc, pg, P)
    assert np.array_equal(got2, got) and (idx2, best2) == (idx, best)
    assert len(set(np.round(want, 12))) > 100
    gm.close(); gsc.close()


@pytest.mark.parametrize("mode", ["obstacle", "max", "mean", "gmapping"])
def test_whole_hill_climbing_match_in_one_launch(sg, gpu, mode):
    """slamgpu_match_hc: every round of HillClimbingScanMatcher::process_scan inside one thread block; pose delta, number of
    poses tested and probability against the oracle's sequential matcher, the log against the oracle's scores"""
    rng = np.random.default_rng(1500)
    model = ob.CELL_GMAPPING if mode == "gmapping" else ob.CELL_TBM_CONSISTENT
    cells = room_map_cells(rng, 240, 240, 0.05, model, passes=4)
    om = ob.OracleMap(240, 240, 0.05, model); om.set_cells(cells)
    gm = sg.GridMap(gpu, 240, 240, 0.05, model); gm.upload(cells)
    kw = dict(obstacle=(ob.OOPE_OBSTACLE, {}), max=(ob.OOPE_MAX, dict(win_v=0.1, win_h=0.1)), mean=(ob.OOPE_MEAN, dict(win_v=0.15, win_h=0.1)),
              gmapping=(ob.OOPE_GMAPPING, dict(gm_th=0.1, gm_window=1)))[mode]
    po, pg = ob.spe_params(kw[0], **kw[1]), sg.spe_params(kw[0], **kw[1])
    for trial in range(4):
        truth = np.array([0.1, -0.2, 0.2]) + rng.normal(0, 0.3, 3) * [1, 1, 0.3]
        r, a = room_scan(rng, 541, 1.5 * np.pi, pose=truth, noise=0.005)
        osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
        init = truth + rng.normal(0, [0.08, 0.08, 0.04])
        limit = (6, 0.1, 0.1) if trial < 3 else (3, 0.07, 0.05)
        m = ob.MatchResult()
        ob.orc.orc_match_hill_climbing(om.h_, C.byref(osc.s), C.byref(po), *init, *limit, C.byref(m), None)
        pose, prob, tested, log = gpu.match_hc(gm, gsc, pg, init, *limit, log_cap=2048)
        st = gpu.score_stats()
        assert st["variant"] == 5, "the match did not take the one-launch path"
        assert tested == m.poses_tested and tested > 20
        assert np.array_equal(pose - init, [m.dx, m.dy, m.dth])
        if mode == "gmapping":
            assert abs(prob - m.best_prob) <= RTOL * abs(m.best_prob)
        else:
            assert prob == m.best_prob
        assert len(log) == tested and np.array_equal(log[0, :3], init)
        want = om.score(osc, po, log[:, :3])
        if mode == "gmapping":
            np.testing.assert_allclose(log[:, 3], want, rtol=RTOL, atol=0)
        else:
            assert np.array_equal(log[:, 3], want)
        # without a log, and with host trig (round-by-round path): same answer
        pose2, prob2, tested2, _ = gpu.match_hc(gm, gsc, pg, init, *limit)
        assert np.array_equal(pose2, pose) and prob2 == prob and tested2 == tested
        ph = sg.spe_params(kw[0], trig=sg.TRIG_HOST, **kw[1])
        pose3, prob3, tested3, log3 = gpu.match_hc(gm, gsc, ph, init, *limit, log_cap=16)
        assert log3 is None and np.array_equal(pose3, pose) and tested3 == tested
        gsc.close()
    gm.close()


def test_measurement_aids(sg, gpu):
    """slamgpu_probe_gather returns a plausible rate; the grid_rows experiment option forces the rows per thread without
    changing a single score"""
    assert gpu.probe_gather(8 << 20, 64) > 1e9
    rng = np.random.default_rng(1800)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, ob.CELL_MEAN, 91)
    xs = p0[0] + 0.02 * (np.arange(60) - 30); ys = p0[1] + 0.02 * (np.arange(50) - 25); ts = p0[2] + 0.01 * (np.arange(6) - 3)
    base, idx0, best0 = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
    for rows in (2, 4, 8):
        gpu.set_option("grid_rows", rows)
        try:
            got, idx, best = gpu.score_grid(gm, gsc, sg.spe_params(), xs, ys, ts)
        finally:
            gpu.set_option("grid_rows", 0)
        assert gpu.score_stats()["rows_per_thread"] == rows
        assert np.array_equal(got, base) and (idx, best) == (idx0, best0)
    with pytest.raises(sg.SlamGpuError):
        gpu.set_option("grid_rows", 3)
    gm.close(); gsc.close()


def test_new_entry_points_on_degenerate_inputs(sg, gpu):
    """an empty scan scores NaN everywhere: hill climbing accepts nothing and still walks its full failed-rounds budget, as
    the oracle's matcher does; an empty chained pose list is a no-op"""
    rng = np.random.default_rng(1900)
    om, gm, osc, gsc, p0 = _setup(sg, gpu, rng, ob.CELL_MEAN, 31)
    empty_g, empty_o = sg.Scan(gpu, np.zeros(0), np.zeros(0)), ob.OracleScan(np.zeros(0), np.zeros(0))
    m = ob.MatchResult()
    op = ob.spe_params()
    ob.orc.orc_match_hill_climbing(om.h_, C.byref(empty_o.s), C.byref(op), *p0, 6, 0.1, 0.1, C.byref(m), None)
    pose, prob, tested, log = gpu.match_hc(gm, empty_g, sg.spe_params(), p0, 6, 0.1, 0.1, log_cap=64)
    assert np.array_equal(pose, p0) and np.isnan(prob) and np.isnan(m.best_prob) and tested == m.poses_tested
    # a budget of zero failed rounds: only the initial pose is scored
    m0 = ob.MatchResult()
    ob.orc.orc_match_hill_climbing(om.h_, C.byref(osc.s), C.byref(op), *p0, 0, 0.1, 0.1, C.byref(m0), None)
    pose, prob, tested, log = gpu.match_hc(gm, gsc, sg.spe_params(), p0, 0, 0.1, 0.1, log_cap=8)
    assert tested == m0.poses_tested == 1 and prob == m0.best_prob and np.array_equal(pose, p0) and len(log) == 1
    scores, states = gpu.score_poses_chained(gm, gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_cache=2), np.zeros((0, 3)))
    assert len(scores) == 0 and len(states) == 0
    gm.close(); gsc.close(); empty_g.close()


@pytest.mark.parametrize("mode", ["obstacle", "mean"])
def test_monte_carlo_segment_in_one_launch(sg, gpu, mode):
    """slamgpu_match_mc: the accept loop of MonteCarloScanMatcher over a given list of pose shifts -- re-basing at every
    accept, the enumerator's counters, the stop at a dispersion reset -- against the same loop run on the host with the
    oracle's scores"""
    rng = np.random.default_rng(1950)
    cells = room_map_cells(rng, 200, 200, 0.05, ob.CELL_MEAN, passes=4)
    om = ob.OracleMap(200, 200, 0.05, ob.CELL_MEAN); om.set_cells(cells)
    gm = sg.GridMap(gpu, 200, 200, 0.05, ob.CELL_MEAN); gm.upload(cells)
    kw = dict(obstacle=(ob.OOPE_OBSTACLE, {}), mean=(ob.OOPE_MEAN, dict(win_v=0.15, win_h=0.1)))[mode]
    po, pg = ob.spe_params(kw[0], **kw[1]), sg.spe_params(kw[0], **kw[1])
    resets = 0
    for trial in range(6):
        truth = np.array([0.1, -0.2, 0.2]) + rng.normal(0, 0.3, 3) * [1, 1, 0.3]
        r, a = room_scan(rng, 360, 2 * np.pi, pose=truth, noise=0.005)
        osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
        init = truth + rng.normal(0, [0.15, 0.15, 0.08])
        max_failed, max_poses = (20, 100) if trial % 2 == 0 else (7, 40)
        noise = rng.normal(0, [0.2, 0.2, 0.1], (max_poses, 3))
        # the reference loop (pose_enumeration_scan_matcher.h:48-65 over GaussianPoseEnumerator::feedback :46-57)
        best, bp = init.copy(), om.score(osc, po, init[None])[0]
        failed = poses = k = 0
        reset = False
        want_log = [np.append(init, bp)]
        while failed < max_failed and poses < max_poses and k < len(noise) and not reset:
            cand = best + noise[k]
            p = om.score(osc, po, cand[None])[0]
            want_log.append(np.append(cand, p))
            k += 1; poses += 1
            if not bp < p:
                failed += 1
                continue
            best, bp = cand, p
            if failed > max_failed // 3:
                failed, reset = 0, True
        res = gpu.match_mc(gm, gsc, pg, init, noise, max_failed, max_poses, log_cap=max_poses + 1)
        assert res is not None and gpu.score_stats()["variant"] == 6
        out, log = res
        assert (out["consumed"], out["failed"], out["poses_nm"], bool(out["reset"])) == (k, failed, poses, reset)
        assert np.array_equal([out["x"], out["y"], out["theta"]], best) and out["prob"] == bp
        assert np.array_equal(log, np.array(want_log))
        resets += reset
        # a second segment continues from the state the first one left
        if reset:
            noise2 = rng.normal(0, [0.1, 0.1, 0.05], (max_poses - poses, 3))
            res2 = gpu.match_mc(gm, gsc, pg, best, noise2, max_failed, max_poses, best_prob=bp, failed=failed, poses_nm=poses, log_cap=max_poses)
            out2, log2 = res2
            assert out2["poses_nm"] >= poses and out2["prob"] >= bp and len(log2) == out2["consumed"]
        gsc.close()
    assert resets > 0, "no trial reached a dispersion reset"
    assert gpu.match_mc(gm, sg.Scan(gpu, r, a), sg.spe_params(kw[0], trig=sg.TRIG_HOST, **kw[1]), init, noise, 20, 100) is None
    gm.close()


def test_whole_match_kernels_at_their_size_limits(sg, gpu):
    """long scans: the one-launch kernels hold 6 (hill climbing) or >= 4 (Monte-Carlo) poses x N terms in shared memory;
    beyond that the hill climbing takes the round-by-round path and the Monte-Carlo entry declines -- same answers"""
    rng = np.random.default_rng(1960)
    cells = room_map_cells(rng, 200, 200, 0.05, ob.CELL_MEAN, passes=3)
    om = ob.OracleMap(200, 200, 0.05, ob.CELL_MEAN); om.set_cells(cells)
    gm = sg.GridMap(gpu, 200, 200, 0.05, ob.CELL_MEAN); gm.upload(cells)
    truth = np.array([0.1, -0.2, 0.2])
    init = truth + [0.06, -0.05, 0.03]
    po, pg = ob.spe_params(), sg.spe_params()
    for n_beams, one_launch in ((4000, True), (4400, False)):
        r, a = room_scan(rng, n_beams, 2 * np.pi, pose=truth, noise=0.005)
        osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
        m = ob.MatchResult()
        ob.orc.orc_match_hill_climbing(om.h_, C.byref(osc.s), C.byref(po), *init, 6, 0.1, 0.1, C.byref(m), None)
        pose, prob, tested, _ = gpu.match_hc(gm, gsc, pg, init, 6, 0.1, 0.1)
        assert (gpu.score_stats()["variant"] == 5) == one_launch
        assert tested == m.poses_tested and prob == m.best_prob and np.array_equal(pose - init, [m.dx, m.dy, m.dth])
        gsc.close()
    noise = rng.normal(0, [0.2, 0.2, 0.1], (30, 3))
    for n_beams, served in ((5000, True), (7000, False)):
        r, a = room_scan(rng, n_beams, 2 * np.pi, pose=truth, noise=0.005)
        gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
        res = gpu.match_mc(gm, gsc, pg, init, noise, 20, 30, log_cap=31)
        assert (res is not None) == served
        if served:
            out, log = res
            assert np.array_equal(log[:, 3], om.score(osc, po, log[:, :3]))
            assert out["prob"] == log[:, 3].max() or out["reset"]
        gsc.close()
    gm.close()
