"""Golden fixtures (tests/golden, generated from the oracle by make_golden.py): the oracle must keep
reproducing them on CPU, and the CUDA path must reproduce them on the GPU without the oracle."""
import json
from pathlib import Path

import numpy as np
import pytest

from dynamicslamtool_b200 import MovingObjectRemoval, Synth
from helpers import IDENTITY_POSE, crc, write_cfg

GOLD = Path(__file__).resolve().parent / "golden"
sys_path_hack = None


def load(name):
    return json.loads((GOLD / name).read_text())


def summarize(m, out):
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", GOLD / "make_golden.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.frame_summary(m, out)


def check(got, want, where):
    for k, v in want.items():
        if k in ("crc_input", "pose"):
            continue
        if k == "centroids":
            a, b = np.array(got[k], np.float64).reshape(-1, 3), np.array(v, np.float64).reshape(-1, 3)
            assert a.shape == b.shape and np.allclose(a, b, rtol=1e-5, atol=1e-9), f"{where}: centroids"
        else:
            assert got[k] == v, f"{where}: {k}: got {got[k]} want {v}"


def run_c1(binding):
    g = load("c1_vlp16_seed1.json")
    m = MovingObjectRemoval(GOLD.parent.parent / g["config"], g["n_bad"], g["n_good"], binding=binding)
    s = Synth(g["scenario"], g["seed"])
    for f, want in enumerate(g["frames"]):
        pts, pose = s.frame(f)
        assert crc(pts) == want["crc_input"], "the seeded generator no longer reproduces the fixture inputs"
        m.push_raw_cloud_and_pose(pts, pose)
        out = m.filter_cloud()
        check(summarize(m, out), want, f"C1 frame {f}")


def run_blobs(binding, tmp_path):
    g = load("two_blobs.json")
    seq = np.load(GOLD / "two_blobs_inputs.npz")["frames"]
    for method in (1, 2):
        cfg = write_cfg(tmp_path, f"m{method}.txt", method_choice=method, **g["overrides"])
        m = MovingObjectRemoval(cfg, g["n_bad"], g["n_good"], binding=binding)
        for f, want in enumerate(g[f"method{method}"]):
            m.push_raw_cloud_and_pose(np.ascontiguousarray(seq[f]), IDENTITY_POSE)
            out = m.filter_cloud()
            check(summarize(m, out), want, f"two_blobs method {method} frame {f}")


def test_oracle_reproduces_golden_c1(oracle):
    run_c1(oracle)


def test_oracle_reproduces_golden_blobs(oracle, tmp_path):
    run_blobs(oracle, tmp_path)


@pytest.mark.gpu
def test_cuda_reproduces_golden_c1(product):
    run_c1(product)


@pytest.mark.gpu
def test_cuda_reproduces_golden_blobs(product, tmp_path):
    run_blobs(product, tmp_path)


# ---------------------------------------------------------------------------------------------- the real reference
# tests/golden/pcl_c1.json is written by oracle/pcl_probe/run_in_docker.sh: the UNMODIFIED reference class on PCL 1.8 over
# the same seeded C1 frames. It cannot be produced in this repository's build image (no ROS, no PCL, no network).
PCL_GOLD = GOLD / "pcl_c1.json"
UNPINNED = ("PARITY UNPINNED: tests/golden/pcl_c1.json is absent - every 'bit-exact' claim of this repository is against its own oracle, "
            "not against a PCL build. Run oracle/pcl_probe/run_in_docker.sh <reference checkout> on a machine with docker to pin it.")


def run_pcl_c1(binding):
    if not PCL_GOLD.exists():
        import warnings
        warnings.warn(UNPINNED)
        pytest.skip(UNPINNED)
    g = json.loads(PCL_GOLD.read_text())
    m = MovingObjectRemoval(GOLD.parent.parent / g["config"], g["n_bad"], g["n_good"], binding=binding)
    s = Synth(g["scenario"], g["seed"])
    for f, want in enumerate(g["frames"]):
        pts, pose = s.frame(f)
        assert crc(pts) == want["crc_input"], "the seeded generator does not reproduce the inputs the reference was run on"
        m.push_raw_cloud_and_pose(pts, pose)
        out = m.filter_cloud()
        check(summarize(m, out), want, f"PCL C1 frame {f}")


def test_oracle_reproduces_pcl_golden(oracle):
    run_pcl_c1(oracle)


@pytest.mark.gpu
def test_cuda_reproduces_pcl_golden(product):
    run_pcl_c1(product)
