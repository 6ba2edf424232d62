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
