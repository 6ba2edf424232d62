"""Parity at the sizes BASELINE.json names (configs[0], [3], [4]); the other GPU tests run the same code at sizes the
oracle finishes in milliseconds.

  configs[0]  tinySLAM: 100x100 @0.1 map, 360 beams, Monte-Carlo matcher (seeded, 0.2 / 0.1, 20 failed, 100 poses), 200 scans
              through the C ABI -- every pose delta, probability, pose count and the map against the oracle's world loop;
  configs[3]  GMapping: 256 particles x 720 beams, per-particle 2560x2560 maps -- batched insertion and the
              one-launch hill climbing; maps / climbed poses of a sample of particles against the oracle;
  configs[4]  4096x4096 @0.025 pyramid, 13 levels, after three inserted scans -- every level bit-equal to the UNMODIFIED
              reference (libslamref.so), and one BF-M3RSM match equal to the reference's matcher.
"""
import ctypes as C

import numpy as np
import pytest

from helpers import room_scan
from oracle import binding as ob

pytestmark = pytest.mark.gpu
RTOL = 1e-5


class _McSampler:
    """GaussianPoseEnumerator's draws (monte_carlo_scan_matcher.h:35-66): mt19937 + three libstdc++ normal distributions,
    restated by the oracle; can be copied so that shifts sampled ahead and not consumed are drawn again"""

    def __init__(self, seed, tr, rot):
        self.g = ob.Mt19937()
        ob.orc.orc_mt_seed(C.byref(self.g), seed)
        self.set_dispersion(tr, rot)

    def set_dispersion(self, tr, rot):
        self.tr, self.rot = tr, rot
        self.n = [ob.Normal(0, tr, 0, 0), ob.Normal(0, tr, 0, 0), ob.Normal(0, rot, 0, 0)]

    def copy(self):
        c = _McSampler.__new__(_McSampler)
        c.g = ob.Mt19937.from_buffer_copy(self.g)
        c.tr, c.rot = self.tr, self.rot
        c.n = [ob.Normal.from_buffer_copy(d) for d in self.n]
        return c

    def draw(self, count):
        out = np.empty((count, 3))
        for k in range(count):
            for a in range(3):
                out[k, a] = ob.orc.orc_normal_sample(C.byref(self.n[a]), C.byref(self.g))
        return out


def _mc_match_gpu(sg, gpu, gm, gsc, params, init, seed, tr, rot, fal, attempts):
    """MonteCarloScanMatcher::process_scan on top of slamgpu_match_mc, the way the C++ plug-in drives it: shifts are sampled
    ahead, the device runs the accept loop until the budget is spent or an accept halves the dispersion"""
    smp = _McSampler(seed, tr, rot)
    best, bp, failed, poses = np.array(init, float), None, 0, 0
    while failed < fal and poses < attempts:
        ahead = smp.copy().draw(attempts - poses)
        res = gpu.match_mc(gm, gsc, params, best, ahead, fal, attempts, best_prob=bp, failed=failed, poses_nm=poses)
        assert res is not None
        out, _ = res
        smp.draw(int(out["consumed"]))  # the real enumerator advances by what was consumed
        best, bp = np.array([out["x"], out["y"], out["theta"]]), out["prob"]
        failed, poses = int(out["failed"]), int(out["poses_nm"])
        if out["reset"]:
            smp.set_dispersion(smp.tr * 0.5, smp.rot * 0.5)
        elif not (failed < fal and poses < attempts):
            break
        else:
            assert out["consumed"] == len(ahead)
    return best - np.array(init, float), bp, poses + 1


def test_configs0_tinyslam_world_200_scans(sg, gpu):
    rng = np.random.default_rng(5000)
    size, scale, n = 100, 0.1, 360
    om = ob.OracleMap(size, size, scale, ob.CELL_MEAN, ob.GROW_PLAIN)
    gm = sg.GridMap(gpu, size, size, scale, sg.CELL_MEAN, sg.GROW_PLAIN)
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    po, pg = ob.spe_params(), sg.spe_params()
    truth = np.array([0.3, -0.2, 0.1])
    est = truth.copy()
    accepted = 0
    for k in range(200):
        step = np.array([0.04 * np.cos(0.05 * k), 0.03 * np.sin(0.07 * k), 0.02 * np.sin(0.11 * k)])
        truth = truth + step
        r, a = room_scan(rng, n, 2 * np.pi, half_w=4.0, half_h=3.0, pose=truth, noise=0.01)
        osc, gsc = ob.OracleScan(r, a), sg.Scan(gpu, r, a)
        init = est + step + rng.normal(0, [0.02, 0.02, 0.01])  # odometry with an error
        if k > 0:
            m = ob.MatchResult()
            ob.orc.orc_match_monte_carlo(om.h_, C.byref(osc.s), C.byref(po), *init, 42 + k, 0.2, 0.1, 20, 100, C.byref(m))
            delta, prob, tested = _mc_match_gpu(sg, gpu, gm, gsc, pg, init, 42 + k, 0.2, 0.1, 20, 100)
            assert np.array_equal(delta, [m.dx, m.dy, m.dth]), k
            assert prob == m.best_prob and tested == m.poses_tested, k
            accepted += bool(np.any(delta != 0))
            est = init + delta
        else:
            est = init
        c1, _ = om.append_scan(osc, est, 1.0, 0, oest, blur=0.5)
        c2 = gpu.append_scan(gm, gsc, est, 1.0, 0, gest, blur=0.5)
        assert c1 == c2, k
        if k % 50 == 49:
            assert gm.info() == om.info()
            assert np.array_equal(gm.download(), om.cells(), equal_nan=True), k
        gsc.close()
    assert accepted > 20 and np.linalg.norm((est - truth)[:2]) < 1.0  # the matcher really moved poses; the track did not diverge
    gm.close()


def test_configs3_gmapping_256_particles_720_beams(sg, gpu):
    rng = np.random.default_rng(5100)
    n, size, scale, beams = 256, 2560, 0.05, 720
    sample = [0, 101, 255]
    parts = sg.Particles(gpu, n, size, size, scale, sg.CELL_GMAPPING, sg.GROW_TILED)
    omaps = {i: ob.OracleMap(size, size, scale, ob.CELL_GMAPPING, ob.GROW_TILED) for i in sample}
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    truth = np.array([1.0, -2.0, 0.3])
    total_cells = 0
    try:
        for k in range(3):
            r, a = room_scan(rng, beams, 2 * np.pi, half_w=17.5, half_h=15.0, pose=truth, noise=0.01)
            poses = truth + rng.normal(0, [0.05, 0.05, 0.01], (n, 3))
            gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
            cells = parts.append_scan(gsc, poses, est=gest)
            total_cells += int(cells.sum())
            for i in sample:
                c, _ = omaps[i].append_scan(osc, poses[i], 1.0, 0, oest)
                assert c == cells[i], (k, i)
            gsc.close()
            truth = truth + [0.1, 0.05, 0.02]
        assert total_cells > 2e8  # ~1e5 cell updates per particle and scan
        for i in sample:
            pm = parts.map(i)
            assert pm.info() == omaps[i].info(), i
            assert np.array_equal(pm.download(), omaps[i].cells(), equal_nan=True), i
        # one-launch hill climbing of every particle against its own map, GMapping OOPE
        r, a = room_scan(rng, beams, 2 * np.pi, half_w=17.5, half_h=15.0, pose=truth, noise=0.005)
        gsc, osc = sg.Scan(gpu, r, a), ob.OracleScan(r, a)
        gparams = sg.spe_params(sg.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
        oparams = ob.spe_params(ob.OOPE_GMAPPING, gm_th=0.1, gm_window=1)
        init = truth + rng.normal(0, [0.04, 0.04, 0.02], (n, 3))
        poses, probs, tested = parts.match_hc(gsc, gparams, init, 6, 0.1, 0.1)
        assert gpu.score_stats()["variant"] == 5  # the whole match of every particle in one launch
        assert tested.min() > 12 and len(set(tested.tolist())) > 3
        for i in sample:
            m = ob.MatchResult()
            ob.orc.orc_match_hill_climbing(omaps[i].h_, C.byref(osc.s), C.byref(oparams), *init[i], 6, 0.1, 0.1, C.byref(m), None)
            assert tested[i] == m.poses_tested, i
            assert np.array_equal(poses[i] - init[i], [m.dx, m.dy, m.dth]), i
            assert abs(probs[i] - m.best_prob) <= RTOL * abs(m.best_prob), i
        gsc.close()
    finally:
        parts.close()


def test_configs4_pyramid_4096_13_levels_and_m3rsm_match(sg, gpu, refso):
    rng = np.random.default_rng(5200)
    size, scale, beams = 4096, 0.025, 1081
    model = ob.CELL_MEAN
    rm = ob.RefMap(size, size, scale, model, ob.GROW_NONE, pyramid_oie=ob.OIE_DISCREPANCY)
    gm = sg.GridMap(gpu, size, size, scale, model, sg.GROW_NONE)
    pyr = sg.Pyramid(gpu, gm, sg.OIE_DISCREPANCY)
    oest, gest = ob.estimator(ob.EST_CONST), sg.estimator(sg.EST_CONST)
    truth = np.array([1.3, -2.1, 0.4])
    try:
        assert pyr.levels() == 13
        for k in range(3):
            pose = truth + rng.normal(0, [0.1, 0.1, 0.05])
            r, a = room_scan(rng, beams, np.deg2rad(270), half_w=30.0, half_h=22.0, pose=pose, noise=0.01)
            occ = np.ones(beams, np.uint8)
            refso.ref_append_scan(rm.h_, beams, ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.u8ptr(occ), pose[0], pose[1], pose[2], 1.0, 0,
                                  C.byref(oest), 0.3, np.inf, 0)
            gsc = sg.Scan(gpu, r, a)
            n2 = pyr.append_scan(gsc, pose, 1.0, 0, gest, blur=0.3)
            gsc.close()
            assert n2 > 1e6
        for lv in range(13):
            assert np.array_equal(pyr.level(lv), rm.export(lv), equal_nan=True), lv
        assert pyr.level(12).shape[:2] == (1, 1)
        r, a = room_scan(rng, beams, np.deg2rad(270), half_w=30.0, half_h=22.0, pose=truth, noise=0.003)
        init = truth + np.array([0.11, -0.07, np.deg2rad(0.8)])
        oparams = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY)
        m = ob.MatchResult()
        occ = np.ones(beams, np.uint8)
        lims = (0.5, 0.5, np.deg2rad(2.0), np.deg2rad(0.5), 0.025)
        refso.ref_match_bf_m3rsm(rm.h_, beams, ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.u8ptr(occ), ob.SPW_EVEN, C.byref(oparams),
                                 *init, *lims, C.byref(m))
        delta, prob, st = pyr.match_m3rsm(r, a, init, sg.spe_params(sg.OOPE_MAX, sg.OIE_DISCREPANCY, prerotated=1), *lims)
        assert np.array_equal(delta, [m.dx, m.dy, m.dth]), (delta, (m.dx, m.dy, m.dth), st)
        assert prob == m.best_prob
        assert st["scored"] > 2 * st["rotations"] and st["branches"] > 3
    finally:
        pyr.close(); gm.close()


def _le(a, b):
    """sg::less_or_equal / math_utils.h:10-51 on arrays"""
    sc = np.maximum(1.0, np.maximum(np.abs(a), np.abs(b)))
    return (np.abs(a - b) <= 1e-7 * sc) | (a < b + 2.220446049250313e-16)


def _fold_level(fine, fi, ci, unknown):
    """one level of RescalableCachingGridMap from the finer one, in numpy: per coarse cell the children in x-outer / y-inner
    order, a child replaces the pick only if its impact is greater by more than eps; MeanProbabilityCell records {p, n},
    discrepancy impact 1 - |p - 1|, unknown = n == 0"""
    H, W = ci["h"], ci["w"]
    Y, X = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    out = np.empty((H, W, 2)); out[...] = unknown
    have = np.zeros((H, W), bool)
    best = np.zeros((H, W))
    for dx in (0, 1):
        for dy in (0, 1):
            sx = 2 * (X - ci["ox"]) + dx + fi["ox"]
            sy = 2 * (Y - ci["oy"]) + dy + fi["oy"]
            ok = (sx >= 0) & (sx < fi["w"]) & (sy >= 0) & (sy < fi["h"])
            rec = fine[np.clip(sy, 0, fi["h"] - 1), np.clip(sx, 0, fi["w"] - 1)]
            known = ok & (rec[..., 1] != 0)
            imp = 1.0 - np.abs(rec[..., 0] - 1.0)
            take = known & (~have | ~_le(imp, best))
            out[take] = rec[take]
            best = np.where(take, imp, best)
            have |= take
    return out


def test_configs4_from_scratch_build_4096(sg, gpu):
    """k_build_fused at the configs[4] size: every level equals the fold of the level below it (bit-equal records), with
    unknown holes and impacts closer than the comparison's eps in the input"""
    rng = np.random.default_rng(5300)
    size = 4096
    cells = np.zeros((size, size, 2))
    cells[..., 0] = np.round(rng.random((size, size)), 6) * 0.999 + rng.integers(0, 3, (size, size)) * 3e-8  # near-ties within eps
    cells[..., 1] = rng.integers(0, 4, (size, size))   # a quarter of the cells unknown
    cells[1000:1400, 2000:2600, 1] = 0                  # and a hole larger than several coarse cells
    cells[..., 0] = np.where(cells[..., 1] == 0, 0.5, cells[..., 0])
    gm = sg.GridMap(gpu, size, size, 0.025, sg.CELL_MEAN, sg.GROW_PLAIN)
    gm.upload(cells)
    pyr = sg.Pyramid(gpu, gm, sg.OIE_DISCREPANCY)
    try:
        pyr.build()
        n = pyr.levels()
        assert n == 13
        fine, fi = cells, pyr.level_info(0)
        unknown = np.array([0.5, 0.0])
        for lv in range(1, n - 1):
            ci = pyr.level_info(lv)
            got = pyr.level(lv)
            want = _fold_level(fine, fi, ci, unknown)
            assert np.array_equal(got, want), lv
            fine, fi = got, ci
        top = pyr.level(n - 1)
        assert top.shape[:2] == (1, 1) and top[0, 0, 1] != 0
    finally:
        pyr.close(); gm.close()
