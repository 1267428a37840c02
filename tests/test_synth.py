import numpy as np

from dynamicslamtool_b200 import Synth


def test_frames_are_a_pure_function_of_seed_and_index(built):
    a, b = Synth(1, 5), Synth(1, 5)
    a.threads, b.threads = 1, 4
    pa, qa = a.frame(3)
    pb, qb = b.frame(3)
    np.testing.assert_array_equal(pa, pb)  # independent of the thread count and of call order
    np.testing.assert_array_equal(qa, qb)
    pc, _ = Synth(1, 6).frame(3)
    assert pc.shape != pa.shape or not np.array_equal(pc, pa)


def test_scenario_shapes(built):
    for scen, lo, hi in ((1, 25_000, 28_800), (2, 110_000, 133_312), (3, 200_000, 262_144), (4, 100_000, 133_312)):
        s = Synth(scen, scen)
        pts, pose = s.frame(0)
        assert lo <= pts.shape[0] <= hi == hi and pts.shape[1] == 4 and pts.dtype == np.float32
        assert s.max_points == hi
        assert np.all(np.isfinite(pts)) and abs(np.linalg.norm(pose[3:]) - 1.0) < 1e-12
