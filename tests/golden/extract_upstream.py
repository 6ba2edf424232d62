#!/usr/bin/env python
"""Extracts the golden vectors the reference's own unit tests hold for the hot
path into JSON fixtures (inputs + expected outputs), so they travel to boxes
without /root/reference.  Run here, once:   python tests/golden/extract_upstream.py

Sources (read-only, parsed with regexes; no reference code is copied):
  test/core/maps/regular_squares_grid_test.cpp         -> upstream_raycast.json
  test/core/geometry_discrete_primitives_test.cpp      -> upstream_bresenham.json
  test/core/maps/area_occupancy_estimator_test.cpp     -> upstream_area_estimator.json
  test/core/scan_matchers/occupancy_observation_probability_test.cpp -> upstream_oope.json
"""
import json
import os
import re
import sys

REF = os.environ.get("SLAM_REF", "/root/reference")
OUT = os.path.dirname(os.path.abspath(__file__))
NUM = r"-?\d+(?:\.\d+)?"


def tests_of(path):
    src = open(os.path.join(REF, path)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    for m in re.finditer(r"TEST(?:_F)?\((\w+),\s*(\w+)\)\s*\{(.*?)\n\}", src, flags=re.S):
        yield m.group(1), m.group(2), m.group(3)


def pts(txt):
    return [[int(a), int(b)] for a, b in re.findall(r"\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}", txt)]


def raycast():
    cases = []
    scale = 0.1  # RSGSegmentRasterizationTest::Grid_Scale, regular_squares_grid_test.cpp:16
    for suite, name, body in tests_of("test/core/maps/regular_squares_grid_test.cpp"):
        if suite != "RSGSegmentRasterizationTest":
            continue
        exp = re.search(r"DSegment\(\{(.*?)\}\)", body, flags=re.S)
        call = re.search(r"world_to_cells\(\{(.*)\}\)", body, flags=re.S).group(1)
        mids = re.findall(r"cell_middle\(\{\s*(-?\d+)\s*,\s*(-?\d+)\s*\}\)", call)
        if mids:
            (ax, ay), (bx, by) = [(int(x), int(y)) for x, y in mids]
            seg = dict(kind="cell_middle", cells=[[ax, ay], [bx, by]])
        else:
            v = [float(x) for x in re.findall(NUM, call)]
            seg = dict(kind="world", pts=[v[0:2], v[2:4]])
        cases.append(dict(name=name, scale=scale, seg=seg, cells=pts(exp.group(1))))
    return cases


def bresenham():
    cases = []
    for suite, name, body in tests_of("test/core/geometry_discrete_primitives_test.cpp"):
        seg = re.search(r"DiscreteSegment2D\{(.*?)\};", body, flags=re.S)
        exp = re.search(r"DPoints\(\{(.*?)\}\)", body, flags=re.S)
        if not seg or not exp:
            continue
        cases.append(dict(name=name, seg=pts(seg.group(1)), cells=pts(exp.group(1))))
    return cases


def area():
    consts = dict(Base_Empty_Prob=0.01, Base_Occup_Prob=0.95, Low_Est_Qual=0.02, Unknown_Est_Qual=0.7)
    src = open(os.path.join(REF, "test/core/maps/area_occupancy_estimator_test.cpp")).read()
    for k in consts:  # area_occupancy_estimator_test.cpp:18-21
        m = re.search(k + r"\s*=\s*(" + NUM + ")", src)
        assert float(m.group(1)) == consts[k]
    cases = []
    for suite, name, body in tests_of("test/core/maps/area_occupancy_estimator_test.cpp"):
        beams = re.findall(r"Segment2D\{\{(.*?)\},\s*\{(.*?)\}\}", body)
        cell = re.search(r"Rectangle\{(.*?)\}", body)
        cellv = [float(x) for x in re.findall(NUM, cell.group(1))] if cell else [-1, 1, -1, 1]
        checks = re.findall(r"ASSERT_EQ\((Occupancy(?:::invalid\(\)|\(.*?\))),\s*aoe\.estimate_occupancy\((\w+),\s*(\w+),\s*(true|false)\)\)",
                            body, flags=re.S)
        if not beams or not checks:
            continue
        named = dict(re.findall(r"auto\s+(\w+)\s*=\s*Segment2D\{(\{.*?\},\s*\{.*?\})\}", body))
        for occ_txt, beam_name, cell_name, is_occ in checks:
            bt = named.get(beam_name)
            if bt is None:
                continue
            bv = [float(x) for x in re.findall(NUM, bt)]
            if "invalid" in occ_txt:
                exp = None
            else:
                args = re.search(r"Occupancy\((.*)\)", occ_txt, flags=re.S).group(1).split(",")
                exp = [consts[a.strip()] if a.strip() in consts else float(eval(a, {'__builtins__': {}}, dict(consts))) for a in args]
            cases.append(dict(name=name, beam=bv, cell=cellv, is_occ=is_occ == "true", expected=exp))
    return dict(base_occupied=[0.95, 1.0], base_empty=[0.01, 1.0], low_qual=0.02, unknown_qual=0.7,
                shift_amount=0.02 * 2.0, cases=cases)


def main():
    if not os.path.isdir(REF):
        sys.exit("reference tree not found: " + REF)
    out = dict(raycast=raycast(), bresenham=bresenham(), area=area())
    for k, v in out.items():
        with open(os.path.join(OUT, "upstream_%s.json" % ("area_estimator" if k == "area" else k)), "w") as f:
            json.dump(v, f, indent=0, separators=(",", ":"))
        n = len(v["cases"]) if isinstance(v, dict) else len(v)
        print(k, n, "cases")


if __name__ == "__main__":
    main()
