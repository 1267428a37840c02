"""GPU parity: CUDA path vs CPU oracle on seeded synthetic sequences, through the C ABI (-m gpu)."""
import numpy as np
import pytest

from dynamicslamtool_b200 import MovingObjectRemoval, Synth
from parity import ParityStats, run_sequence

pytestmark = pytest.mark.gpu


def _frames(scenario, seed, n, start=0):
    s = Synth(scenario, seed)
    for f in range(start, start + n):
        yield s.frame(f)


@pytest.mark.parametrize("scenario,cfg,frames,method", [
    (1, "MOR_config.txt", 40, 2),          # C1: VLP-16, reference default config
    (2, "MOR_config_hdl64.txt", 14, 2),    # C2: HDL-64E, 120k points
    (1, "MOR_config.txt", 14, 1),          # method 1 (point distance estimate)
])
def test_sequence_parity(product, oracle, cfg_dir, tmp_path, scenario, cfg, frames, method):
    path = cfg_dir / cfg
    if method != 2:
        text = path.read_text().replace("method_choice:2", f"method_choice:{method}")
        path = tmp_path / cfg
        path.write_text(text)
    gpu = MovingObjectRemoval(path, 4, 3, binding=product)
    orc = MovingObjectRemoval(path, 4, 3, binding=oracle)
    stats = ParityStats()
    first_bad, bad = run_sequence(gpu, orc, _frames(scenario, scenario, frames), stats)
    print("parity stats", stats.as_dict())
    assert first_bad is None, f"first divergence at frame {first_bad}: {bad}"
    assert stats.matches > 0
    if scenario == 1 and method == 2:
        assert stats.removed_points > 0 and stats.mo_frames > 0  # the removal path was exercised
