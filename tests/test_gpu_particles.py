"""K6 parity on the GPU: the GMapping particle step batched over particles, against the oracle run particle by
particle.  Obstacle-OOPE runs are bit-exact (scores, poses, number of poses tested); the GMapping OOPE goes
through exp(), so its scores are compared at rtol 1e-5 and the climbed poses must still coincide."""
import ctypes as C

import numpy as np
import pytest

from helpers import room_scan
from oracle import binding as ob

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _build(sg, gpu, rng, n, model, size=160, scale=0.05, scans=3):
    """n particles with slightly different pose histories -> n different maps, mirrored in the oracle"""
    parts = sg.Particles(gpu, n, size, size, scale, model, sg.GROW_TILED)
    omaps = [ob.OracleMap(size, size, scale, model, ob.GROW_TILED) for _ in range(n)]
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    truth = np.array([0.2, -0.1, 0.3])
    for k in range(scans):
        r, a = room_scan(rng, 181, 2 * np.pi, half_w=3.0, half_h=2.5, pose=truth, noise=0.01)
        poses = truth + rng.normal(0, [0.03, 0.03, 0.01], (n, 3))
        gsc = sg.Scan(gpu, r, a)
        upd = np.ones(n, np.uint8)
        if k == 1:
            upd[::3] = 0  # some particles skip an update (scan probability 0 upstream)
        cells = parts.append_scan(gsc, poses, do_update=upd, est=gest)
        for i in range(n):
            if upd[i]:
                c, _ = omaps[i].append_scan(ob.OracleScan(r, a), poses[i], 1.0, 0, oest)
                assert c == cells[i]
            else:
                assert cells[i] == 0
        gsc.close()
        truth = truth + [0.05, 0.02, 0.02]
    return parts, omaps, truth


def _check_maps(parts, omaps):
    for i, om in enumerate(omaps):
        pm = parts.map(i)
        assert pm.info() == om.info(), i
        assert np.array_equal(pm.download(), om.cells(), equal_nan=True), i


@pytest.mark.parametrize("model", [ob.CELL_MEAN, ob.CELL_GMAPPING])
def test_particle_maps_scores_and_hill_climbing(sg, gpu, model):
    rng = np.random.default_rng(4000 + model)
    n = 12
    parts, omaps, truth = _build(sg, gpu, rng, n, model)
    _check_maps(parts, omaps)
    r, a = room_scan(rng, 121, 2 * np.pi, half_w=3.0, half_h=2.5, pose=truth, noise=0.005)
    gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
    gm_mode = model == ob.CELL_GMAPPING
    if gm_mode:
        gparams = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
        oparams = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    else:
        gparams, oparams = sg.spe_params(), ob.spe_params()
    # per-particle candidate scoring, one launch
    cand = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 7, 3))
    got = parts.score(gsc, gparams, cand)
    want = np.stack([omaps[i].score(osc, oparams, cand[i]) for i in range(n)])
    if gm_mode:
        np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    else:
        assert np.array_equal(got, want)
    assert len(set(np.round(want[:, 0], 12))) > 1  # the maps really differ
    # lock-step hill climbing
    init = truth + rng.normal(0, [0.04, 0.04, 0.02], (n, 3))
    active = np.ones(n, np.uint8); active[5] = 0
    poses, probs, tested = parts.match_hc(gsc, gparams, init, 6, 0.1, 0.1, active=active)
    for i in range(n):
        if not active[i]:
            assert tested[i] == 0 and np.array_equal(poses[i], init[i])
            continue
        m = ob.MatchResult()
        ob.orc.orc_match_hill_climbing(omaps[i].h_, C.byref(osc.s), C.byref(oparams), *init[i], 6, 0.1, 0.1, C.byref(m), None)
        assert tested[i] == m.poses_tested, i
        assert np.array_equal(poses[i] - init[i], [m.dx, m.dy, m.dth]), i
        if gm_mode:
            assert abs(probs[i] - m.best_prob) <= RTOL * abs(m.best_prob)
        else:
            assert probs[i] == m.best_prob
    assert tested[active > 0].min() > 30
    gsc.close(); parts.close()


def test_resample_copies_maps_on_device(sg, gpu):
    rng = np.random.default_rng(4100)
    n = 8
    parts, omaps, truth = _build(sg, gpu, rng, n, ob.CELL_GMAPPING, size=96, scans=2)
    src = np.array([3, 3, 2, 7, 3, 5, 0, 7], np.int32)   # multinomial draw: duplicates, a kept-in-place particle, dropped ones
    parts.resample(src)
    after = [omaps[s] for s in src]
    _check_maps(parts, after)
    # copies are independent: updating one does not touch its siblings
    r, a = room_scan(rng, 91, 2 * np.pi, half_w=3.0, half_h=2.5, pose=truth)
    gsc = sg.Scan(gpu, r, a)
    upd = np.zeros(n, np.uint8); upd[1] = 1
    parts.append_scan(gsc, np.tile(truth, (n, 1)), do_update=upd)
    clone = ob.OracleMap(model=ob.CELL_GMAPPING, handle=ob.orc.orc_map_clone(omaps[3].h_), owner=True)
    clone.append_scan(ob.OracleScan(r, a), truth, 1.0, 0, ob.estimator())
    after[1] = clone
    _check_maps(parts, after)
    gsc.close(); parts.close()


@pytest.mark.parametrize("grow", ["tiled", "plain"])
def test_batched_insertion_grows_each_map_like_the_oracle(sg, gpu, grow):
    """every particle's map starts too small, so the batched insertion replays a different growth per map"""
    rng = np.random.default_rng(4100)
    n = 9
    g_gpu = sg.GROW_TILED if grow == "tiled" else sg.GROW_PLAIN
    g_or = ob.GROW_TILED if grow == "tiled" else ob.GROW_PLAIN
    parts = sg.Particles(gpu, n, 40, 40, 0.1, ob.CELL_TBM_CONSISTENT, g_gpu)
    omaps = [ob.OracleMap(40, 40, 0.1, ob.CELL_TBM_CONSISTENT, g_or) for _ in range(n)]
    oest, gest = ob.estimator(ob.EST_AREA), sg.estimator(sg.EST_AREA)
    truth = np.array([0.3, 0.1, -0.2])
    for k in range(3):
        r, a = room_scan(rng, 151, 2 * np.pi, half_w=4.0 + k, half_h=3.0 + k, pose=truth, noise=0.01)
        poses = truth + rng.normal(0, [0.2, 0.2, 0.05], (n, 3))
        gsc = sg.Scan(gpu, r, a)
        cells = parts.append_scan(gsc, poses, est=gest, blur=0.3)
        for i in range(n):
            c, _ = omaps[i].append_scan(ob.OracleScan(r, a), poses[i], 1.0, 0, oest, blur=0.3)
            assert c == cells[i]
        gsc.close()
        truth = truth + [0.4, -0.3, 0.1]
    infos = {tuple(sorted(parts.map(i).info().items())) for i in range(n)}
    assert len(infos) > 1 or grow == "tiled"  # the maps really diverged (tiled growth may coincide)
    _check_maps(parts, omaps)


def test_hill_climbing_with_the_estimators_lifetime_cache(sg, gpu):
    """GmappingOccupancyObservationPE caches (cell -> probability) for its whole life: with gm_cache=2 every particle
    carries its own estimator's cache through the rounds of one match and on into the next scan, like the oracle run
    particle by particle with one cache object per particle"""
    rng = np.random.default_rng(4200)
    n = 8
    parts, omaps, truth = _build(sg, gpu, rng, n, ob.CELL_GMAPPING)
    gparams = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2)
    oparams = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
    caches = [ob.GmCache(0, 0, -1.0) for _ in range(n)]
    for k in range(3):  # the cache survives from scan to scan
        r, a = room_scan(rng, 360, 2 * np.pi, half_w=3.0, half_h=2.5, pose=truth, noise=0.005)
        gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
        init = truth + rng.normal(0, [0.04, 0.04, 0.02], (n, 3))
        active = np.ones(n, np.uint8); active[k] = 0
        poses, probs, tested = parts.match_hc(gsc, gparams, init, 6, 0.1, 0.1, active=active)
        assert gpu.score_stats()["variant"] == 5  # the whole match of every particle in one launch, cache included
        for i in range(n):
            if not active[i]:
                continue
            m = ob.MatchResult()
            ob.orc.orc_match_hill_climbing(omaps[i].h_, C.byref(osc.s), C.byref(oparams), *init[i], 6, 0.1, 0.1, C.byref(m),
                                           C.byref(caches[i]))
            assert tested[i] == m.poses_tested, (k, i)
            assert np.array_equal(poses[i] - init[i], [m.dx, m.dy, m.dth]), (k, i)
            assert abs(probs[i] - m.best_prob) <= 1e-12 * abs(m.best_prob)
        gsc.close()
        truth = truth + [0.02, 0.01, 0.01]
    parts.close()


def test_carried_cache_round_by_round_path_agrees(sg, gpu):
    """host trig forces the lock-step path (one launch per round): same poses, same counts, same cache afterwards"""
    rng = np.random.default_rng(4300)
    n = 5
    parts_a, omaps, truth = _build(sg, gpu, rng, n, ob.CELL_GMAPPING, scans=2)
    rng = np.random.default_rng(4300)
    parts_b, _, _ = _build(sg, gpu, rng, n, ob.CELL_GMAPPING, scans=2)
    for k in range(2):
        r, a = room_scan(rng, 360, 2 * np.pi, half_w=3.0, half_h=2.5, pose=truth, noise=0.005)
        gsc = sg.Scan(gpu, r, a)
        init = truth + rng.normal(0, [0.04, 0.04, 0.02], (n, 3))
        fast = parts_a.match_hc(gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2), init, 6, 0.1, 0.1)
        v_fast = gpu.score_stats()["variant"]
        slow = parts_b.match_hc(gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1, gm_cache=2, trig=sg.TRIG_HOST), init, 6, 0.1, 0.1)
        assert v_fast == 5 and gpu.score_stats()["variant"] != 5
        assert np.array_equal(fast[0], slow[0]) and np.array_equal(fast[2], slow[2])
        np.testing.assert_allclose(fast[1], slow[1], rtol=1e-12, atol=0)
        gsc.close()
    parts_a.close(); parts_b.close()


def test_copy_on_write_tiles_of_the_particle_maps(sg, gpu):
    """LazyTiledGridMap (lazy_tiled_grid_map.h:18-118): a new particle map references one shared all-unknown tile, a scan
    makes only the tiles under it private, resampling hands out tile references instead of copying cells, and the first
    insertion after it clones exactly the tiles it writes.  Every map stays bit-equal to the oracle's throughout."""
    rng = np.random.default_rng(4400)
    n, size, scale = 16, 1024, 0.05                       # 8 x 8 tiles of 128 x 128 cells per map
    parts = sg.Particles(gpu, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
    omaps = [ob.OracleMap(size, size, scale, ob.CELL_GMAPPING, ob.GROW_TILED) for _ in range(n)]
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    st = parts.tile_stats()
    assert st["tiled"] == 1 and st["tiles_live"] == 1     # only the shared unknown tile exists
    truth = np.array([0.3, -0.2, 0.1])

    def insert(upd=None):
        r, a = room_scan(rng, 240, 2 * np.pi, half_w=6.0, half_h=5.0, pose=truth, noise=0.01)   # 12 x 10 m room: 2 x 2 .. 3 x 3 tiles
        poses = truth + rng.normal(0, [0.03, 0.03, 0.01], (n, 3))
        gsc = sg.Scan(gpu, r, a)
        cells = parts.append_scan(gsc, poses, do_update=upd, est=gest)
        for i in range(n):
            if upd is None or upd[i]:
                c, _ = omaps[i].append_scan(ob.OracleScan(r, a), poses[i], 1.0, 0, oest)
                assert c == cells[i]
        gsc.close()

    insert()
    st1 = parts.tile_stats()
    per_map = (st1["tiles_live"] - 1) / n
    assert 4 <= per_map <= 16 and st1["tiles_live"] - 1 == st1["tiles_cloned"]   # a few of the 64 tiles per map, all cloned from tile 0
    assert st1["pool_bytes"] < 0.3 * n * size * size * 5 * 8                  # far below 16 dense maps
    _check_maps(parts, omaps)
    insert()                                                                   # same room: nothing new to clone
    assert parts.tile_stats()["tiles_cloned"] == st1["tiles_cloned"]
    # resample: particle i becomes a copy of src[i]; 0 and 5 survive, the others are copies of them
    src = np.array([0, 0, 0, 5, 5, 5, 5, 0, 0, 5, 5, 0, 0, 5, 0, 5], np.int32)
    parts.resample(src)
    st2 = parts.tile_stats()
    assert st2["resample_bytes"] == 0 and st2["resample_tiles_shared"] > 0    # table copies only
    assert st2["tiles_live"] < st1["tiles_live"]                               # the dropped particles' tiles went back to the pool
    after = []
    for i in range(n):
        after.append(omaps[src[i]] if src[i] == i else
                     ob.OracleMap(model=ob.CELL_GMAPPING, handle=ob.orc.orc_map_clone(omaps[src[i]].h_), owner=True))
    omaps = after
    _check_maps(parts, omaps)
    # an insertion into SOME of the copies clones only their tiles; the untouched copies keep sharing
    upd = np.zeros(n, np.uint8); upd[[1, 2, 3]] = 1
    live_before = parts.tile_stats()["tiles_live"]
    insert(upd)
    grown = parts.tile_stats()["tiles_live"] - live_before
    assert 3 * 4 <= grown <= 3 * 16
    _check_maps(parts, omaps)
    insert()
    _check_maps(parts, omaps)
    # scoring reads through the tile tables
    r, a = room_scan(rng, 181, 2 * np.pi, half_w=6.0, half_h=5.0, pose=truth, noise=0.005)
    gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
    cand = truth + rng.normal(0, [0.05, 0.05, 0.02], (n, 5, 3))
    got = parts.score(gsc, sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1), cand)
    want = np.stack([omaps[i].score(osc, ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1), cand[i]) for i in range(n)])
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=0)
    gsc.close(); parts.close()
