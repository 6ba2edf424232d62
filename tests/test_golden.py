"""The oracle against the golden vectors the reference's own unit tests hold for the hot path
(tests/golden/upstream_*.json, extracted from /root/reference/test by tests/golden/extract_upstream.py).
CPU only; this is the pin that travels to boxes without /root/reference."""
import ctypes as C
import json
import os

import numpy as np
import pytest

from oracle import binding as ob

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    return json.load(open(os.path.join(G, name)))


def segment_of(case):
    sc = case["scale"]
    if case["seg"]["kind"] == "cell_middle":  # RegularSquaresGrid::cell_to_world, regular_squares_grid.h:103-105
        (ax, ay), (bx, by) = case["seg"]["cells"]
        return [sc * (ax + 0.5), sc * (ay + 0.5), sc * (bx + 0.5), sc * (by + 0.5)]
    (ax, ay), (bx, by) = case["seg"]["pts"]
    return [ax, ay, bx, by]


def occupancy_equal(got, exp):
    """Occupancy::operator== (src/core/states/state_data.h:17-21): both invalid, or are_equal on both fields"""
    if exp is None:
        return bool(np.isnan(got).any())
    return all(ob.orc.orc_are_equal(float(g), float(e)) for g, e in zip(got, exp))


def area_estimator(a):
    return ob.estimator(ob.EST_AREA, occ=tuple(a["base_occupied"]), empty=tuple(a["base_empty"]), low_qual=a["low_qual"],
                        unknown_qual=a["unknown_qual"], shift=a["shift_amount"])


def test_raycast_golden_vectors():
    cases = load("upstream_raycast.json")
    assert len(cases) == 47
    for c in cases:
        got = ob.raycast(ob.orc.orc_raycast, *segment_of(c), c["scale"])
        assert got.tolist() == c["cells"], c["name"]


def test_bresenham_golden_vectors():
    cases = load("upstream_bresenham.json")
    assert len(cases) == 24
    buf = np.zeros((256, 2), np.int32)
    for c in cases:
        (ax, ay), (bx, by) = c["seg"]
        n = ob.orc.orc_bresenham(ax, ay, bx, by, ob.iptr(buf), 256)
        assert buf[:n].tolist() == c["cells"], c["name"]


def test_area_estimator_golden_vectors():
    a = load("upstream_area_estimator.json")
    assert len(a["cases"]) == 64
    est = area_estimator(a)
    out = np.zeros(2)
    for c in a["cases"]:
        ob.orc.orc_estimate_occupancy(C.byref(est), *c["beam"], *[float(v) for v in c["cell"]], int(c["is_occ"]), ob.dptr(out))
        assert occupancy_equal(out, c["expected"]), (c["name"], out.tolist(), c["expected"])


@pytest.mark.gpu
def test_gpu_raycast_golden_vectors(sg, gpu):
    cases = load("upstream_raycast.json")
    by_scale = {}
    for c in cases:
        by_scale.setdefault(c["scale"], []).append(c)
    for sc, cs in by_scale.items():
        offs, cells = gpu.raycast_segments(sc, [segment_of(c) for c in cs])
        for i, c in enumerate(cs):
            assert cells[offs[i]:offs[i + 1]].tolist() == c["cells"], c["name"]


@pytest.mark.gpu
def test_gpu_area_estimator_golden_vectors(sg, gpu):
    a = load("upstream_area_estimator.json")
    est = sg.estimator(sg.EST_AREA, occ=tuple(a["base_occupied"]), empty=tuple(a["base_empty"]), low_qual=a["low_qual"],
                       unknown_qual=a["unknown_qual"], shift=a["shift_amount"])
    cs = a["cases"]
    out = gpu.estimate_occupancy(est, [c["beam"] for c in cs], [c["cell"] for c in cs], [c["is_occ"] for c in cs])
    oest = area_estimator(a)
    want = np.zeros(2)
    for c, got in zip(cs, out):
        assert occupancy_equal(got, c["expected"]), (c["name"], got.tolist(), c["expected"])
        ob.orc.orc_estimate_occupancy(C.byref(oest), *c["beam"], *[float(v) for v in c["cell"]], int(c["is_occ"]), ob.dptr(want))
        assert np.array_equal(got, want, equal_nan=True), c["name"]  # and bit-equal to the oracle
