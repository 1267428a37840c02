"""The ROS-free C++ harness (harness/mov_harness.cpp) drives the C++ class MovingObjectRemoval exactly like the
reference node drives the reference class; its per-frame output CRCs must equal the oracle's."""
import subprocess

import pytest

from dynamicslamtool_b200 import MovingObjectRemoval, Synth
from helpers import ROOT, crc

pytestmark = pytest.mark.gpu


def test_cpp_class_matches_oracle(oracle, built):
    exe = ROOT / "harness" / "mov_harness"
    if not exe.exists():
        subprocess.check_call(["make", "-C", str(ROOT / "harness")])
    cfg = ROOT / "config" / "MOR_config.txt"
    n = 10
    res = subprocess.run([str(exe), str(cfg), "1", "1", str(n)], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stderr
    lines = [l.split() for l in res.stdout.splitlines() if l.startswith("frame ")]
    assert len(lines) == n
    orc = MovingObjectRemoval(cfg, 4, 3, binding=oracle)
    s = Synth(1, 1)
    for f in range(n):
        pts, pose = s.frame(f)
        orc.push_raw_cloud_and_pose(pts, pose)
        out = orc.filter_cloud()
        # frame f in N out M crc X ms T
        assert int(lines[f][3]) == pts.shape[0] and int(lines[f][5]) == out.shape[0]
        assert int(lines[f][7], 16) == crc(out), f"frame {f}: C++ class output differs from the oracle"


def test_harness_config_error_is_an_exception_not_exit0(built):
    exe = ROOT / "harness" / "mov_harness"
    res = subprocess.run([str(exe), "/nonexistent.txt", "1", "1", "1"], capture_output=True, text=True, timeout=60)
    assert res.returncode != 0  # the reference would print "Couldnt open the file" and exit(0)
