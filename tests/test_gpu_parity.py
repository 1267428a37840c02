"""GPU parity: CUDA path vs CPU oracle on seeded synthetic sequences, through the C ABI (-m gpu)."""
import numpy as np
import pytest

from dynamicslamtool_b200 import MovingObjectRemoval, Synth
from parity import ParityStats, compare_frame, run_sequence

pytestmark = pytest.mark.gpu


def _frames(scenario, seed, n, start=0):  # noqa: D401
    s = Synth(scenario, seed)
    for f in range(start, start + n):
        yield s.frame(f)


@pytest.mark.parametrize("scenario,cfg,frames,method,start", [
    (1, "MOR_config.txt", 40, 2, 0),          # C1: VLP-16, reference default config
    (2, "MOR_config_hdl64.txt", 14, 2, 0),    # C2: HDL-64E, 120k points
    (2, "MOR_config_hdl64.txt", 7, 2, 101),   # C2 where a mover passes at 1.7 m: cells with > 1000 points, 27k-point clusters
    (3, "MOR_config_os128.txt", 7, 2, 0),     # C3: 128 beams, 262k points, ~100 clusters, size ties
    (1, "MOR_config.txt", 14, 1, 0),          # method 1 (point distance estimate)
    (2, "MOR_config_hdl64.txt", 6, 1, 0),     # method 1 at outdoor scale
])
def test_sequence_parity(product, oracle, cfg_dir, tmp_path, scenario, cfg, frames, method, start):
    path = cfg_dir / cfg
    if method != 2:
        text = path.read_text().replace("method_choice:2", f"method_choice:{method}")
        path = tmp_path / cfg
        path.write_text(text)
    gpu = MovingObjectRemoval(path, 4, 3, binding=product, max_points=Synth(scenario, scenario).max_points)
    orc = MovingObjectRemoval(path, 4, 3, binding=oracle)
    stats = ParityStats()
    first_bad, bad = run_sequence(gpu, orc, _frames(scenario, scenario, frames, start), stats)
    print("parity stats", stats.as_dict())
    assert first_bad is None, f"first divergence at frame {first_bad}: {bad}"
    assert stats.matches > 0
    if scenario == 1 and method == 2:
        assert stats.removed_points > 0 and stats.mo_frames > 0  # the removal path was exercised


@pytest.mark.parametrize("scenario,base,frames,overrides", [
    (4, "MOR_config_terrain.txt", 8, {}),                                                  # C4: eigen-normal mode on sloped / multi-plane terrain
    (4, "MOR_config_terrain.txt", 6, {"ground_mode": 1, "gp_leaf": 0.5}),                  # literal voxel-covariance test, outdoor leaf
    (1, "MOR_config.txt", 8, {"ground_mode": 1}),                                          # literal mode with the reference defaults (leaf 0.1, bin_gap 10)
    (1, "MOR_config.txt", 6, {"ground_mode": 2, "gp_bin_width": 0.1}),
    (4, "MOR_config_terrain.txt", 3, {"gp_leaf": 0.1}),                                    # a voxel grid of ~10 million cells (> 9000 scan tiles)
])
def test_voxel_covariance_ground_modes(product, oracle, cfg_dir, tmp_path, scenario, base, frames, overrides):
    """Ground removal by voxel covariance (reference cpp:90-200, dead code there): parity against the oracle's repaired
    restatement, including the per-voxel taps (centroids / normals within 1e-5, accepted flags and bins exact)."""
    import numpy as np
    from helpers import write_cfg
    cfg = write_cfg(tmp_path, base=cfg_dir / base, **overrides)
    maxp = Synth(scenario, scenario).max_points
    gpu = MovingObjectRemoval(cfg, 4, 3, binding=product, max_points=maxp)
    orc = MovingObjectRemoval(cfg, 4, 3, binding=oracle)
    stats = ParityStats()
    for f, (pts, pose) in enumerate(_frames(scenario, scenario, frames)):
        gpu.push_raw_cloud_and_pose(pts, pose)
        orc.push_raw_cloud_and_pose(pts, pose)
        vg, vo = gpu.tap("ground_voxels"), orc.tap("ground_voxels")
        assert vg.shape == vo.shape and vg.shape[0] > 0, f"frame {f}: voxel count {vg.shape} vs {vo.shape}"
        np.testing.assert_allclose(vg[:, :3], vo[:, :3], rtol=1e-5, atol=1e-7, err_msg=f"frame {f}: voxel centroids")
        assert np.array_equal(vg[:, 3], vo[:, 3]), f"frame {f}: accepted flags differ in {int(np.sum(vg[:, 3] != vo[:, 3]))} voxels"
        acc = vo[:, 3] > 0
        assert np.array_equal(vg[acc, 4], vo[acc, 4]), f"frame {f}: bin keys"
        np.testing.assert_allclose(vg[acc, 5:], vo[acc, 5:], rtol=1e-5, atol=1e-6, err_msg=f"frame {f}: normals")
        og, oo = gpu.filter_cloud().copy(), orc.filter_cloud().copy()
        bad = compare_frame(gpu, orc, og, oo, stats)
        assert not bad, f"frame {f}: {bad}"
        assert gpu.counts()["NG"] > 0
    print("ground parity stats", stats.as_dict())
