"""GPU parity: CUDA path vs CPU oracle on seeded synthetic sequences, through the C ABI (-m gpu)."""
import numpy as np
import pytest

from dynamicslamtool_b200 import MovingObjectRemoval, Synth
from parity import ParityStats, run_sequence

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
