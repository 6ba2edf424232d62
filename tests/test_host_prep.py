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
