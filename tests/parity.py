"""Frame-by-frame comparison of the CUDA path against the CPU oracle through the shared tap ABI.

Bar (BASELINE.json north_star): bit-exact for point classes, component labels, cluster membership,
cluster order, matches, moving flags, mo_vec confidences, removal mask and the output cloud bytes;
1e-5 relative for centroids and quantities derived from them (reported: how many are not bit-equal).
"""
from __future__ import annotations

import numpy as np

REL_TOL = 1e-5  # north_star: "within 1e-5 relative for transformed coordinates, voxel centroids and normals"

EXACT_COUNTS = ["N", "NT", "NC", "NG", "K", "KPREV", "M", "MU", "NMO", "NOUT", "NKPREV", "P1", "P2", "TWO_FRAMES",
                "EXTRACT_OVERFLOW", "NCPREV", "NK"]


def _close(a, b):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    if a.shape != b.shape:
        return False
    nan = np.isnan(a)
    if not np.array_equal(nan, np.isnan(b)):
        return False
    a, b = a[~nan], b[~nan]
    return bool(np.all(np.abs(a - b) <= REL_TOL * np.maximum(np.abs(a), np.abs(b)) + 1e-12))


class ParityStats:
    def __init__(self):
        self.frames = 0
        self.centroid_values = 0
        self.centroid_not_bitexact = 0
        self.removed_points = 0
        self.mo_frames = 0
        self.flagged = 0
        self.matches = 0
        self.overflow_frames = 0
        self.size_tie_frames = 0

    def as_dict(self):
        return dict(self.__dict__)


def compare_frame(gpu, orc, out_gpu, out_orc, stats: ParityStats | None = None, after_filter=True):
    """Returns a list of human-readable mismatches (empty = parity)."""
    bad = []
    cg, co = gpu.counts(), orc.counts()
    for k in EXACT_COUNTS:
        if not after_filter and k in ("NOUT", "EXTRACT_OVERFLOW"):
            continue
        if cg[k] != co[k]:
            bad.append(f"count {k}: gpu {cg[k]} oracle {co[k]}")
    if cg["ERRFLAGS"]:
        bad.append(f"gpu error flags {cg['ERRFLAGS']}")
    if bad:
        return bad  # shapes differ; the rest would be noise

    def exact(name):
        a, b = gpu.tap(name), orc.tap(name)
        if a.shape != b.shape or not np.array_equal(a, b):
            n = int(np.sum(a != b)) if a.shape == b.shape else -1
            bad.append(f"{name}: {n} of {a.size} differ (bit-exact required)")

    def close(name, count_bits=False):
        a, b = gpu.tap(name), orc.tap(name)
        if not _close(a, b):
            bad.append(f"{name}: outside {REL_TOL} relative")
        elif count_bits and stats is not None:
            stats.centroid_values += a.size
            stats.centroid_not_bitexact += int(np.sum(a.view(np.uint32) != b.view(np.uint32)))

    for name in ("point_class", "labels", "cluster_id", "cluster_root", "cluster_size", "flags"):
        exact(name)
    close("centroids", count_bits=True)
    close("cluster_bbox")
    if co["TWO_FRAMES"]:
        exact("transform")
        close("prev_centroids_t")
        a, b = gpu.tap("prev_points_t"), orc.tap("prev_points_t")
        if not (a.shape == b.shape and np.array_equal(a.view(np.uint32), b.view(np.uint32))):
            # transformed coordinates: tolerance per north_star, but the arithmetic is replicated exactly
            if not _close(a, b):
                bad.append("prev_points_t: outside tolerance")
            else:
                bad.append(f"prev_points_t: within tolerance but {int(np.sum(a.view(np.uint32) != b.view(np.uint32)))} values not bit-equal")
        close("prev_bbox_t")
        for name in ("recip_query", "recip_match", "match_query", "match_match", "match_score"):
            exact(name)
        close("match_dist")
    exact("mo_conf")
    close("mo_centroids")
    if after_filter:
        exact("removed_mask")
        exact("cluster_removed")
        if out_gpu.shape != out_orc.shape or not np.array_equal(out_gpu.view(np.uint32), out_orc.view(np.uint32)):
            bad.append(f"output cloud bytes differ: gpu {out_gpu.shape} oracle {out_orc.shape}")
    if stats is not None:
        stats.frames += 1
        stats.matches += co["M"]
        stats.flagged += int(orc.tap("flags").sum())
        stats.mo_frames += 1 if co["NMO"] else 0
        stats.overflow_frames += co["EXTRACT_OVERFLOW"]
        stats.size_tie_frames += 1 if co["SIZE_TIES"] else 0
        if after_filter:
            stats.removed_points += int(np.sum(orc.tap("removed_mask") == 2))
    return bad


def run_sequence(gpu, orc, frames, stats: ParityStats | None = None, stop_on_first=True):
    """frames: iterable of (points[n,4] float32, pose7). Returns (first_bad_frame, mismatches)."""
    for f, (pts, pose) in enumerate(frames):
        gpu.push_raw_cloud_and_pose(pts, pose)
        orc.push_raw_cloud_and_pose(pts, pose)
        og = gpu.filter_cloud().copy()
        oo = orc.filter_cloud().copy()
        bad = compare_frame(gpu, orc, og, oo, stats)
        if bad and stop_on_first:
            return f, bad
    return None, []
