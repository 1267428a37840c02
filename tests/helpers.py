"""Shared helpers for the CPU and GPU test suites."""
from __future__ import annotations

import zlib
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
DEFAULT_CFG = ROOT / "config" / "MOR_config.txt"


def write_cfg(tmp_path, name="cfg.txt", base=DEFAULT_CFG, drop=(), extra=(), **overrides):
    """Copy a MOR_config.txt with some values replaced / keys dropped / raw lines appended."""
    lines = []
    for line in Path(base).read_text().splitlines():
        key = line.split(":", 1)[0]
        if key in drop:
            continue
        if key in overrides:
            line = f"{key}:{overrides.pop(key)}"
        lines.append(line)
    for k, v in overrides.items():
        lines.append(f"{k}:{v}")
    lines.extend(extra)
    p = Path(tmp_path) / name
    p.write_text("\n".join(lines) + "\n")
    return p


OPEN_CFG = dict(trim_x=1000.0, trim_y=1000.0, trim_z=1000.0, gp_limit=-1000.0)  # nothing trimmed, nothing cropped

IDENTITY_POSE = np.array([0, 0, 0, 0, 0, 0, 1], np.float64)


def f32_sqdist_matrix(p):
    """FLANN L2_Simple<float> for all pairs, evaluated in float32 exactly like the oracle / kernels."""
    p = np.asarray(p, np.float32)
    dx = p[:, None, 0] - p[None, :, 0]
    dy = p[:, None, 1] - p[None, :, 1]
    dz = p[:, None, 2] - p[None, :, 2]
    return (dx * dx + dy * dy) + dz * dz


def components_min_label(adj):
    """Connected components of a dense boolean adjacency; label = min index of the component."""
    from scipy.sparse import csr_matrix
    from scipy.sparse.csgraph import connected_components
    n = adj.shape[0]
    _, lab = connected_components(csr_matrix(adj), directed=False)
    mins = np.full(lab.max() + 1, n, np.int64)
    np.minimum.at(mins, lab, np.arange(n))
    return mins[lab].astype(np.int32)


def crc(a) -> int:
    return zlib.crc32(np.ascontiguousarray(a).tobytes()) & 0xFFFFFFFF


def blob(rng, center, n, sigma=0.05):
    return (np.asarray(center, np.float32) + rng.normal(0, sigma, (n, 3))).astype(np.float32)


def with_intensity(xyz, value=0.5):
    return np.ascontiguousarray(np.concatenate([xyz, np.full((len(xyz), 1), value, np.float32)], axis=1), np.float32)
