"""Drop-in parity: slam-constructor's own SingleStateHypothesisLaserScanGridWorld (the tinySLAM / vinySLAM
world class) run twice on the same synthetic scans -- once with the reference's CPU scan matcher, map and
scan adder, once with the CUDA plug-ins of slam_constructor_b200/host/slamgpu_backend.h -- inside one
process.  The binary is built against the unmodified reference headers by host/Makefile (in the container
that has /root/reference) and travels to the GPU box; it checks poses, observer callbacks and every map
cell for equality and prints RESULT: ALL PASSED."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "slam_constructor_b200", "host", "_build", "test_dropin")
BIN_KEY = os.path.join(ROOT, "slam_constructor_b200", "host", "_build", "test_backend_key")
PATCH = os.path.join(ROOT, "integration", "slam_backend_key.patch")
REF = "/root/reference"


def _run():
    return subprocess.run([BIN], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900)


def test_dropin_binary_is_built_and_refuses_cpu(sg):
    """CPU box: the adapters compile against the reference headers, link against libslamgpu.so, and the
    binary stops at context creation (exit 77) because there is no CPU fallback"""
    if not os.path.exists(BIN):
        pytest.skip("host/_build/test_dropin not built (needs /root/reference at build time)")
    if sg.lib().slamgpu_device_count() > 0:
        pytest.skip("a B200 is visible")
    r = _run()
    assert r.returncode == 77 and "NO-DEVICE" in r.stdout, r.stdout


@pytest.mark.gpu
def test_reference_worlds_with_cuda_plugins_match_cpu_reference():
    assert os.path.exists(BIN), "host/_build/test_dropin missing: run __graft_entry__.build() where /root/reference exists"
    r = _run()
    print(r.stdout)
    assert r.returncode == 0 and "RESULT: ALL PASSED" in r.stdout, r.stdout[-4000:]


def test_backend_key_patch_applies_to_the_reference_factories(tmp_path):
    """integration/slam_backend_key.patch (`slam/backend=cuda` inside init_scan_matcher / init_grid_map / init_scan_adder)
    applies cleanly to a throw-away copy of the reference's two factory headers; it adds 8 lines and changes one"""
    if not os.path.isdir(os.path.join(REF, "src", "utils")):
        pytest.skip("the reference tree is not on this box")
    dst = tmp_path / "src" / "utils"
    dst.mkdir(parents=True)
    for f in ("init_scan_matching.h", "init_occupancy_mapping.h"):
        (dst / f).write_text(open(os.path.join(REF, "src", "utils", f)).read())
    r = subprocess.run(["patch", "-p1", "-d", str(tmp_path), "-i", PATCH], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0, r.stdout
    txt = (dst / "init_scan_matching.h").read_text() + (dst / "init_occupancy_mapping.h").read_text()
    assert txt.count("slamgpu_hooks::wants_cuda(props)") == 3 and txt.count('#include "slamgpu_factory_hooks.h"') == 2
    added = [ln for ln in open(PATCH) if ln.startswith("+") and not ln.startswith("+++")]
    assert len(added) == 6


def test_backend_key_binary_is_built_and_refuses_cpu(sg):
    if not os.path.exists(BIN_KEY):
        pytest.skip("host/_build/test_backend_key not built (needs /root/reference at build time)")
    if sg.lib().slamgpu_device_count() > 0:
        pytest.skip("a B200 is visible")
    r = subprocess.run([BIN_KEY], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    assert r.returncode == 77 and "NO-DEVICE" in r.stdout, r.stdout


@pytest.mark.gpu
def test_reference_factory_selects_the_cuda_backend_by_preset_key():
    """the reference's own init_1h_slam(props), compiled from the patched headers: `slam/backend=cuda` in the preset gives a
    world on the CUDA plug-ins whose poses and map equal the CPU world's built by the same factory from the same preset"""
    assert os.path.exists(BIN_KEY), "host/_build/test_backend_key missing: run __graft_entry__.build() where /root/reference exists"
    r = subprocess.run([BIN_KEY], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600)
    print(r.stdout)
    assert r.returncode == 0 and "RESULT: ALL PASSED" in r.stdout, r.stdout[-4000:]
