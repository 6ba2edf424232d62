"""The per-scan host preparation of the library (scan filter, point weights, mapping quality: O(n) libm code, no
GPU needed) against the oracle.  Runs on the CPU box: these entry points never touch the device."""
import ctypes as C

import numpy as np
import pytest

from oracle import binding as ob


@pytest.mark.parametrize("kind", [0, 1, 2])
def test_point_weights_match_oracle(sg, kind):
    rng = np.random.default_rng(50 + kind)
    for n in (1, 2, 57, 360):
        r = rng.uniform(0.3, 8, n); a = np.sort(rng.uniform(-2.3, 2.3, n))
        want = np.empty(n)
        ob.orc.orc_point_weights(kind, n, ob.dptr(r), ob.dptr(a), ob.dptr(want))
        assert np.array_equal(sg.point_weights(kind, r, a), want)


def test_mapping_quality_matches_oracle(sg):
    rng = np.random.default_rng(60)
    r = rng.uniform(0.3, 8, 200); a = np.linspace(-2, 2, 200)
    v = np.zeros(200, np.uint32)
    ob.orc.orc_angle_histogram_values(200, ob.dptr(r), ob.dptr(a), v.ctypes.data_as(C.POINTER(C.c_uint32)))
    assert np.array_equal(sg.mapping_quality(sg.OMQE_AHR, r, a), 1.0 / v)
    assert np.array_equal(sg.mapping_quality(sg.OMQE_IDLE, r, a), np.ones(200))


@pytest.mark.parametrize("budget", [(6, 0.1, 0.1), (3, 0.07, 0.05), (0, 0.1, 0.1), (9, 0.2, 0.02)])
def test_hill_climbing_state_machine_against_the_oracle_matcher(sg, budget):
    """hill_climb.h -- the enumerator + accept loop shared by the one-launch device kernel and the round-by-round path --
    driven on the host by the oracle's scoring function: candidates, accept decisions and pose count must be the oracle
    matcher's (restatement of HillClimbingScanMatcher, pinned against the reference build in test_oracle_pin.py)"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import room_map_cells, room_scan
    rng = np.random.default_rng(70 + budget[0])
    cells = room_map_cells(rng, 160, 160, 0.05, ob.CELL_MEAN, passes=3)
    om = ob.OracleMap(160, 160, 0.05, ob.CELL_MEAN)
    om.set_cells(cells)
    params = ob.spe_params()
    L = sg.lib()
    cb_t = C.CFUNCTYPE(C.c_double, C.POINTER(C.c_double), C.c_void_p)
    L.slamgpu_debug_hill_climb.argtypes = [C.POINTER(C.c_double), C.c_uint32, C.c_double, C.c_double, cb_t, C.c_void_p,
                                           C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    for trial in range(3):
        truth = np.array([0.1, -0.2, 0.2]) + rng.normal(0, 0.2, 3) * [1, 1, 0.3]
        r, a = room_scan(rng, 181, 1.5 * np.pi, pose=truth, noise=0.005)
        osc = ob.OracleScan(r, a)
        init = truth + rng.normal(0, [0.08, 0.08, 0.04])
        calls = []

        def score(pose, _user):
            p = np.array([pose[0], pose[1], pose[2]])
            calls.append(p)
            return float(om.score(osc, params, p[None])[0])

        m = ob.MatchResult()
        ob.orc.orc_match_hill_climbing(om.h_, C.byref(osc.s), C.byref(params), *init, *budget, C.byref(m), None)
        out, prob, tested = np.zeros(3), C.c_double(), C.c_int64()
        ini = np.ascontiguousarray(init, dtype=np.float64)
        rc = L.slamgpu_debug_hill_climb(ini.ctypes.data_as(C.POINTER(C.c_double)), budget[0], budget[1], budget[2], cb_t(score), None,
                                        out.ctypes.data_as(C.POINTER(C.c_double)), C.byref(prob), C.byref(tested))
        assert rc == 0
        assert tested.value == m.poses_tested == len(calls)
        assert np.array_equal(out - init, [m.dx, m.dy, m.dth]) and prob.value == m.best_prob


@pytest.mark.parametrize("seed,lims", [(80, (0.4, 0.4, 3.0, 0.5, 0.05)), (81, (0.6, 0.3, 2.0, 1.0, 0.1))])
def test_m3rsm_engine_against_the_reference_matcher(sg, refso, seed, lims):
    """the host branch-and-bound engine of slamgpu_match_m3rsm (queue order, pruning, branching, the speculative filling of
    every scoring call), run on the CPU over Match bounds computed by the unmodified reference: it must return what the
    reference's BruteForceMultiResolutionScanMatcher returns on the same map and scan"""
    import sys, os
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from helpers import room_scan
    rng = np.random.default_rng(seed)
    rm = ob.RefMap(256, 256, 0.05, ob.CELL_MEAN, ob.GROW_PLAIN, pyramid_oie=ob.OIE_DISCREPANCY)
    est = ob.estimator(ob.EST_CONST)
    truth = np.array([0.3, -0.2, 0.1])
    for k in range(3):
        pose = truth + rng.normal(0, [0.1, 0.1, 0.05])
        r, a = room_scan(rng, 200, 2 * np.pi, pose=pose)
        refso.ref_append_scan(rm.h_, 200, ob.dptr(ob.f64(r)), ob.dptr(ob.f64(a)), ob.u8ptr(np.ones(200, np.uint8)), pose[0], pose[1], pose[2],
                              1.0, 0, C.byref(est), 0.3, np.inf, 0)
    xlim, ylim, rot_deg, ang_deg, tstep = lims
    r, a = room_scan(rng, 90, np.deg2rad(270), pose=truth, noise=0.003)
    r, a = ob.f64(r), ob.f64(a)
    init = truth + np.array([0.11, -0.07, np.deg2rad(0.8)])
    m = ob.MatchResult()
    oparams = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY)
    refso.ref_match_bf_m3rsm(rm.h_, len(r), ob.dptr(r), ob.dptr(a), ob.u8ptr(np.ones(len(r), np.uint8)), ob.SPW_EVEN, C.byref(oparams), *init,
                             xlim, ylim, np.deg2rad(rot_deg), np.deg2rad(ang_deg), tstep, C.byref(m))
    bparams = ob.spe_params(ob.OOPE_MAX, ob.OIE_DISCREPANCY, prerotated=1)
    asked = [0]

    def bounds(count, rots, wins, out, _user):
        asked[0] += count
        for k in range(count):
            rot = rots[k]
            # LaserScan2D::to_cartesian(rot + pose.theta), as slamgpu_match_m3rsm pre-rotates the scan
            x, y = ob.f64(0 + r * np.cos(rot + init[2] + a)), ob.f64(0 + r * np.sin(rot + init[2] + a))
            out[k] = refso.ref_match_bound(rm.h_, len(r), ob.dptr(x), ob.dptr(y), 1, ob.SPW_EVEN, C.byref(bparams), init[0], init[1], init[2],
                                           rot, wins[4 * k], wins[4 * k + 1], wins[4 * k + 2], wins[4 * k + 3])

    cb_t = C.CFUNCTYPE(None, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double), C.c_void_p)
    L = sg.lib()
    L.slamgpu_debug_m3rsm.argtypes = [C.c_double] * 6 + [cb_t, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    delta, prob, st = np.zeros(3), C.c_double(), np.zeros(4, np.int64)
    rc = L.slamgpu_debug_m3rsm(xlim, ylim, np.deg2rad(rot_deg), np.deg2rad(ang_deg), tstep, 0.0, cb_t(bounds), None,
                               delta.ctypes.data_as(C.POINTER(C.c_double)), C.byref(prob), st.ctypes.data_as(C.POINTER(C.c_int64)))
    assert rc == 0
    assert np.array_equal(delta, [m.dx, m.dy, m.dth]), (delta, (m.dx, m.dy, m.dth))
    assert prob.value == m.best_prob
    assert st[0] == asked[0] and st[1] < st[2] + 2  # far fewer scoring calls than branches: the calls were filled up
