"""sm_runner's on-disk triple (pose2D / scan2D / map): files written by the
unmodified reference (tests/golden/gen_sm_runner_fixture.cpp) are parsed, round-tripped byte for byte, and on
the GPU the hill-climbing matcher replays them to the reference's own answer."""
import filecmp
import json
import os

import numpy as np
import pytest

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sm_runner")
CASES = [("affine_cell", 1), ("tbm_cell", 4)]  # name, SLAMGPU_CELL_* model


@pytest.mark.parametrize("name,model", CASES)
def test_fixture_round_trip(sg, name, model, tmp_path):
    from slam_constructor_b200 import fixtures
    exp = json.load(open(os.path.join(G, "expected.json")))[name]
    m = fixtures.read_map(os.path.join(G, name + ".map"), model)
    assert (m["w"], m["h"], m["ox"], m["oy"]) == (exp["width"], exp["height"], exp["origin"][0], exp["origin"][1])
    known = m["cells"][..., {1: 1, 4: 5}[model]]
    assert 500 < known.sum() < m["w"] * m["h"]
    fixtures.write_map(tmp_path / "m.map", m["cells"], model, m["scale"], m["ox"], m["oy"])
    assert filecmp.cmp(tmp_path / "m.map", os.path.join(G, name + ".map"), shallow=False)
    r, a, occ = fixtures.read_scan2d(os.path.join(G, name + ".scan2D"))
    assert len(r) == 91 and occ.sum() < 91
    fixtures.write_scan2d(tmp_path / "s.scan2D", r, a, occ)
    r2, a2, occ2 = fixtures.read_scan2d(tmp_path / "s.scan2D")
    assert np.array_equal(r, r2) and np.array_equal(a, a2) and np.array_equal(occ, occ2)
    pose = fixtures.read_pose2d(os.path.join(G, name + ".pose2D"))
    fixtures.write_pose2d(tmp_path / "p.pose2D", pose)
    assert np.array_equal(fixtures.read_pose2d(tmp_path / "p.pose2D"), pose)


@pytest.mark.gpu
@pytest.mark.parametrize("name,model", CASES)
def test_replay_dumped_state_on_gpu(sg, gpu, name, model):
    from slam_constructor_b200 import fixtures
    exp = json.load(open(os.path.join(G, "expected.json")))[name]["hc"]
    m = fixtures.read_map(os.path.join(G, name + ".map"), model)
    r, a, occ = fixtures.read_scan2d(os.path.join(G, name + ".scan2D"))
    pose = fixtures.read_pose2d(os.path.join(G, name + ".pose2D"))
    parts = sg.Particles(gpu, 1, m["w"], m["h"], m["scale"], model, sg.GROW_PLAIN)
    gm = parts.map(0)
    gm.upload(m["cells"], m["ox"], m["oy"])
    keep = gm.filter_scan(r, a, pose, occ=occ)                      # the SPE's filter_scan, then its even weights
    fr, fa = r[keep], a[keep]
    scan = sg.Scan(gpu, fr, fa, weight=sg.point_weights(sg.SPW_EVEN, fr, fa))
    poses, probs, tested = parts.match_hc(scan, sg.spe_params(), pose[None], 6, 0.1, 0.1)
    assert probs[0] == exp["prob"]
    assert np.array_equal(poses[0] - pose, exp["delta"])
    scan.close(); parts.close()


@pytest.mark.parametrize("name,model", CASES)
def test_pgm_export_matches_reference_dumper(sg, name, model, tmp_path):
    """row f4: the PGM the reference's GridMapToPgmDumber wrote for the same map, byte for byte"""
    from slam_constructor_b200 import fixtures
    m = fixtures.read_map(os.path.join(G, name + ".map"), model)
    fixtures.write_pgm(tmp_path / "m.pgm", m["cells"])
    assert filecmp.cmp(tmp_path / "m.pgm", os.path.join(G, name + ".pgm"), shallow=False)
