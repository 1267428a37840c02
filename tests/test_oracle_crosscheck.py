"""CPU tests that pin the oracle against independent implementations available offline
(numpy float32 brute force, scipy connected components / Rotation) and hand-derived timelines.

The reference ships no golden vectors (parity unpinned, see oracle/mor_oracle.cpp); these tests
are what keeps the restatement honest.
"""
import ctypes as C

import numpy as np
import pytest

from dynamicslamtool_b200 import MovingObjectRemoval
from helpers import (IDENTITY_POSE, OPEN_CFG, blob, components_min_label, f32_sqdist_matrix, with_intensity, write_cfg)


def make(oracle, tmp_path, n_bad=4, n_good=3, **kw):
    return MovingObjectRemoval(write_cfg(tmp_path, **kw), n_bad, n_good, binding=oracle)


# ------------------------------------------------------------------ trim + crop (A1, A2)
def test_trim_and_crop_classes(oracle, tmp_path):
    m = make(oracle, tmp_path)  # defaults: trim 3/3/5, gp_limit -0.5
    pts = np.array([
        [0, 0, 0, 1], [3.0, 0, 0, 1], [3.0000002, 0, 0, 1], [-3.0, -3.0, -0.5, 1], [0, 0, -0.50000006, 1],
        [0, 0, 5.0, 1], [0, 0, 5.000001, 1], [np.nan, 0, 0, 1], [0, np.inf, 0, 1], [0, 0, -np.inf, 1], [0, -3.1, 0, 1],
    ], np.float32)
    m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
    cls = m.tap("point_class")
    #               in   edge  out  corner  below  top  above nan inf  -inf  y-out
    assert list(cls) == [1, 1, 0, 1, 2, 1, 2, 0, 0, 0, 0]
    out = m.filter_cloud()
    # output = cloud points in order, then the gp_indices points in order (cpp:673-684); wire layout x,y,z,1,i,0,0,0
    assert out.shape == (6, 8)
    np.testing.assert_array_equal(out[:, :3], pts[[0, 1, 3, 5, 4, 6], :3])
    assert np.all(out[:, 3] == 1.0) and np.all(out[:, 4] == 1.0) and np.all(out[:, 5:] == 0)
    assert list(m.tap("removed_mask")) == [1, 1, 0, 1, 1, 1, 1, 0, 0, 0, 0]


def test_empty_and_tiny_frames(oracle, tmp_path):
    m = make(oracle, tmp_path)
    for n in (0, 1, 0, 3):
        m.push_raw_cloud_and_pose(np.zeros((n, 4), np.float32), IDENTITY_POSE)
        assert m.filter_cloud().shape[0] == n
        assert m.counts()["K"] == 0


# ------------------------------------------------------------------ Euclidean clustering (A5-A10)
@pytest.mark.parametrize("seed,n,r", [(0, 1500, 0.11), (1, 2500, 0.3), (2, 800, 0.05)])
def test_clustering_matches_bruteforce_float32(oracle, tmp_path, seed, n, r):
    rng = np.random.default_rng(seed)
    # clumpy cloud: uniform background + blobs, so components of every size exist
    xyz = np.concatenate([rng.uniform(-2, 2, (n // 2, 3)), *(blob(rng, rng.uniform(-2, 2, 3), n // 10, 0.08) for _ in range(5))]).astype(np.float32)
    m = make(oracle, tmp_path, ec_distance_threshold=r, min_cluster_size=5, max_cluster_size=400, **OPEN_CFG)
    m.push_raw_cloud_and_pose(with_intensity(xyz), IDENTITY_POSE)
    r2 = np.float32(np.float64(np.float32(r)) * np.float64(np.float32(r)))  # A7
    adj = f32_sqdist_matrix(xyz) < r2                                       # A6: strict
    want = components_min_label(adj)
    np.testing.assert_array_equal(m.tap("labels"), want)
    # size filter + order (size desc, min index asc) + per-cluster centroid (double sums in index order)
    roots, sizes = np.unique(want, return_counts=True)
    keep = (sizes >= 5) & (sizes <= 400)
    order = np.lexsort((roots[keep], -sizes[keep]))
    np.testing.assert_array_equal(m.tap("cluster_root"), roots[keep][order])
    np.testing.assert_array_equal(m.tap("cluster_size"), sizes[keep][order])
    cent = m.tap("centroids")
    for k, root in enumerate(roots[keep][order]):
        members = np.flatnonzero(want == root)
        acc = np.zeros(3, np.float64)
        for i in members:  # sequential double accumulation, ascending index (compute3DCentroid)
            acc += xyz[i].astype(np.float64)
        np.testing.assert_array_equal(cent[k], (acc / len(members)).astype(np.float32))
    cid = m.tap("cluster_id")
    assert np.all((cid >= 0) == np.isin(want, roots[keep]))


def test_clustering_radius_is_strict_and_float(oracle, tmp_path):
    # two points exactly at the float radius apart are NOT neighbours (d2 < r2 strict)
    r = np.float32(0.25)
    m = make(oracle, tmp_path, ec_distance_threshold=0.25, min_cluster_size=1, max_cluster_size=10, **OPEN_CFG)
    pts = with_intensity(np.array([[0, 0, 0], [r, 0, 0], [10, 0, 0], [10 + np.nextafter(r, np.float32(0)), 0, 0]], np.float32))
    m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
    d = f32_sqdist_matrix(pts[:, :3])
    r2 = np.float32(np.float64(r) * np.float64(r))
    np.testing.assert_array_equal(m.tap("labels"), components_min_label(d < r2))
    assert list(m.tap("labels")[:2]) == [0, 1]


# ------------------------------------------------------------------ pose delta (A11, A12)
def test_pose_delta_against_scipy(oracle):
    from scipy.spatial.transform import Rotation as R
    rng = np.random.default_rng(3)
    fn = oracle.lib.oracle_pose_delta
    fn.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_float)]
    for _ in range(50):
        qa, qb = R.random(random_state=rng.integers(1 << 30)), R.random(random_state=rng.integers(1 << 30))
        ta, tb = rng.uniform(-50, 50, 3), rng.uniform(-50, 50, 3)
        prev = np.concatenate([ta, qa.as_quat()])
        cur = np.concatenate([tb, qb.as_quat()])
        out = (C.c_float * 12)()
        fn((C.c_double * 7)(*prev), (C.c_double * 7)(*cur), out)
        M = np.array(out[:], np.float64).reshape(3, 4)
        Rrel = qb.inv() * qa
        np.testing.assert_allclose(M[:, :3], Rrel.as_matrix(), atol=3e-7)
        np.testing.assert_allclose(M[:, 3], qb.inv().apply(ta - tb), rtol=1e-6, atol=1e-6)
    # identical poses => exact identity
    out = (C.c_float * 12)()
    p = (C.c_double * 7)(1.5, -2.0, 0.3, 0.1, 0.2, 0.3, 0.9273618495495703)
    fn(p, p, out)
    np.testing.assert_allclose(np.array(out[:]).reshape(3, 4), np.eye(3, 4), atol=1e-7)


def test_unnormalised_quaternion_is_scaled_not_normalised(oracle):
    # tf::Matrix3x3::setRotation uses s = 2/|q|^2: a scaled quaternion gives the same rotation (A11)
    fn = oracle.lib.oracle_pose_delta
    fn.argtypes = [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_float)]
    a, b = (C.c_float * 12)(), (C.c_float * 12)()
    ident = (C.c_double * 7)(0, 0, 0, 0, 0, 0, 1)
    fn((C.c_double * 7)(1, 2, 3, 0.1, 0.2, 0.3, 0.9), ident, a)
    fn((C.c_double * 7)(1, 2, 3, 0.2, 0.4, 0.6, 1.8), ident, b)
    np.testing.assert_allclose(np.array(a[:]), np.array(b[:]), atol=1e-7)


# ------------------------------------------------------------------ matching + moving tests + tracking timeline
def two_blob_frames(rng, n_frames, step, n=400):
    """A static blob at (2,0,0) and a blob moving +x by `step` per frame from (-2,1,0); identity odometry."""
    base_s = blob(rng, (2, 0, 0), n, 0.12)
    base_m = blob(rng, (-2, 1, 0), n, 0.12)
    for f in range(n_frames):
        jitter = rng.normal(0, 0.002, (2 * n, 3)).astype(np.float32)
        yield with_intensity(np.concatenate([base_s, base_m + np.float32([step * f, 0, 0])]) + jitter)


def octree_new_voxel_count(c1, c2, res=np.float64(np.float32(0.1))):
    """Independent restatement of OctreePointCloudChangeDetector (A13 as refined in DESIGN.md)."""
    eps = np.float64(np.finfo(np.float32).eps)
    first = c1[0].astype(np.float64)
    lo, hi = first - res / 2, first + res / 2
    mn = lo - ((2.0 * res - eps) - (hi - lo)) / 2.0
    k1 = {tuple(v) for v in np.floor((c1.astype(np.float64) - mn) / res).astype(np.int64)}
    return sum(tuple(v) not in k1 for v in np.floor((c2.astype(np.float64) - mn) / res).astype(np.int64))


def test_method2_scores_and_confirmation_timeline(oracle, tmp_path):
    rng = np.random.default_rng(7)
    m = make(oracle, tmp_path, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    frames = list(two_blob_frames(rng, 9, 0.12))
    prev = None
    nmo = []
    for f, pts in enumerate(frames):
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        c = m.counts()
        assert c["K"] == 2
        if f > 0:
            assert c["M"] == 2 and c["TWO_FRAMES"] == 1
            # scores = octree new-voxel counts of the matched pairs, recomputed independently
            cid_prev, cid_cur = prev[1], m.tap("cluster_id")
            for q, mm, score in zip(m.tap("match_query"), m.tap("match_match"), m.tap("match_score")):
                c1 = prev[0][cid_prev == q, :3]
                c2 = pts[cid_cur == mm, :3]
                assert score == octree_new_voxel_count(c1, c2)
                thr = (len(c1) + len(c2)) // 20                      # cpp:590 unsigned integer division
                assert bool(m.tap("flags")[mm]) == (score > thr)
        out = m.filter_cloud()
        nmo.append(m.counts()["NMO"])
        prev = (pts, m.tap("cluster_id").copy())
        if nmo[-1]:
            removed = m.tap("removed_mask") == 2
            assert removed.sum() == 400 and out.shape[0] == 400
            assert np.all(pts[removed, 0] < 1.0)                     # the moving blob is the one removed
    # n_bad = 4: flags exist from frame 1, four consecutive flagged frames are 1..4 => confirmed in frame 4's push
    assert nmo[:4] == [0, 0, 0, 0] and nmo[4] == 1 and all(v == 1 for v in nmo[4:])
    assert list(m.tap("mo_conf")) == [4]                             # n_good + 1, capped (.h:91-93)


def test_tracked_mover_decays_when_it_stops(oracle, tmp_path):
    rng = np.random.default_rng(8)
    m = make(oracle, tmp_path, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    moving = list(two_blob_frames(rng, 6, 0.12))
    still = [moving[-1].copy() for _ in range(8)]
    conf = []
    for pts in moving + still:
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        m.filter_cloud()
        conf.append(list(m.tap("mo_conf")))
    assert conf[4] == [4] and conf[5] == [4]
    # once static (identical clouds => 0 new voxels => flag false) confidence drops by 1 per frame and the entry is erased at 0
    assert conf[6:10] == [[3], [2], [1], []]
    assert conf[-1] == []


def test_method1_scores(oracle, tmp_path):
    rng = np.random.default_rng(9)
    m = make(oracle, tmp_path, method_choice=1, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    frames = list(two_blob_frames(rng, 3, 0.12))
    prev = None
    for f, pts in enumerate(frames):
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        if f > 0:
            cid_prev, cid_cur = prev[1], m.tap("cluster_id")
            for q, mm, score in zip(m.tap("match_query"), m.tap("match_match"), m.tap("match_score")):
                c1, c2 = prev[0][cid_prev == q, :3], pts[cid_cur == mm, :3]
                d = f32_sqdist_matrix(np.concatenate([c1, c2]))[: len(c1), len(c1):].min(axis=1)
                count = np.sum((d > np.float32(0.005)) & (d < np.float32(0.5)))   # squared distance vs pde bounds, cpp:356
                assert score == count / ((len(c1) + len(c2)) // 2)
                assert bool(m.tap("flags")[mm]) == (score > float(np.float32(0.15)))
        m.filter_cloud()
        prev = (pts, m.tap("cluster_id").copy())


def test_volume_constraint_rejects_mismatched_boxes(oracle, tmp_path):
    rng = np.random.default_rng(10)
    m = make(oracle, tmp_path, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    a = blob(rng, (0, 0, 0), 500, 0.1)
    m.push_raw_cloud_and_pose(with_intensity(a), IDENTITY_POSE)
    m.filter_cloud()
    b = (a * np.float32([2.2, 1, 1])).astype(np.float32)  # same centroid region, volume x2.2 => |dv|/(sum) = 0.375 > 0.3
    m.push_raw_cloud_and_pose(with_intensity(b), IDENTITY_POSE)
    c = m.counts()
    assert c["MU"] == 1 and c["M"] == 0


def test_extract_overflow_quirk(oracle, tmp_path):
    """A18: when mo_vec selects the same cluster more often than `cloud` has points, ExtractIndices fails and the
    non-ground part of the output is empty. Reproduced by confirming a mover, then teleporting it by > catch_up_distance
    (a second entry is pushed) while both entries still resolve to the same, only, cluster."""
    rng = np.random.default_rng(11)
    m = make(oracle, tmp_path, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    base = blob(rng, (0, 0, 0), 300, 0.1)
    seen_overflow = False
    for f in range(12):
        pts = with_intensity(base + np.float32([0.45 * f, 0, 0]))
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        out = m.filter_cloud()
        c = m.counts()
        if c["EXTRACT_OVERFLOW"]:
            seen_overflow = True
            assert c["NMO"] >= 2 and out.shape[0] == 0 and np.all(m.tap("removed_mask") == 2)
    assert seen_overflow


# ------------------------------------------------------------------ VISUALIZE outputs and reset
def test_cluster_collection_markers_and_reset(oracle, tmp_path):
    """cluster_collection (cpp:226-229) = the clustered points, cluster after cluster, ascending cloud index inside a
    cluster - rebuilt here from the cluster_id tap; markers (cpp:7-58, :640-642): float centroid, extents with the
    0 -> 0.1 rule, ids 1, 2, 3, ... in mo_vec order (the counter of cpp:622 advances at cpp:669); oracle_reset = a newly constructed object."""
    rng = np.random.default_rng(7)
    m = make(oracle, tmp_path, ec_distance_threshold=0.3, min_cluster_size=50, max_cluster_size=5000, **OPEN_CFG)
    frames = list(two_blob_frames(rng, 9, 0.12))
    sig = []
    for pts in frames:
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        cid = m.tap("cluster_id")                                   # OPEN_CFG: every input point is a cloud point
        coll = m.cluster_collection()
        want = np.concatenate([pts[cid == k] for k in range(m.counts()["K"])])
        assert coll.shape == (want.shape[0], 8)
        assert np.array_equal(coll[:, :3], want[:, :3]) and np.array_equal(coll[:, 4], want[:, 3])
        assert np.all(coll[:, 3] == 1.0) and not coll[:, 5:].any()
        n_tracked = m.counts()["NMO"]
        m.filter_cloud()
        marks = m.moving_markers()
        assert marks.shape[0] == n_tracked
        for i, mk in enumerate(marks):
            p = pts[cid == mk["cluster"], :3]
            ext = p.max(0) - p.min(0)
            assert np.array_equal(mk["scale"], np.where(ext == 0, np.float32(0.1), ext))
            np.testing.assert_allclose(mk["position"], p.astype(np.float64).mean(0), rtol=1e-5, atol=1e-5)
            assert mk["id"] == i + 1 and np.allclose(mk["color"], [0.8, 0.1, 0.4, 0.5])
        sig.append((m.counts()["NMO"], m.tap("removed_mask").tobytes()))
    assert any(s[0] for s in sig)
    m.reset()
    sig2 = []
    for pts in frames:
        m.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        m.filter_cloud()
        sig2.append((m.counts()["NMO"], m.tap("removed_mask").tobytes()))
    assert sig2 == sig


# ------------------------------------------------------------------ order independence
@pytest.mark.parametrize("seed", [21, 22, 23])
def test_partition_and_output_do_not_depend_on_point_order(oracle, tmp_path, seed):
    """EuclideanClusterExtraction grows whole components before the size test (A5), so the partition of the cloud into
    clusters - and with it the set of removed points - cannot depend on the order of the points in the message. Labels
    (minimum index) and the order inside clusters do; the sets must not."""
    rng = np.random.default_rng(seed)
    kw = dict(ec_distance_threshold=0.3, min_cluster_size=40, max_cluster_size=5000, **OPEN_CFG)
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    a, b = make(oracle, tmp_path / "a", **kw), make(oracle, tmp_path / "b", **kw)
    static = [blob(rng, (rng.uniform(-8, 8), rng.uniform(-8, 8), 0), int(rng.integers(60, 300)), 0.15) for _ in range(4)]
    mover = blob(rng, (0, -5, 0), 250, 0.12)
    small = blob(rng, (6, 6, 2), 20, 0.05)                      # below min_cluster_size
    for f in range(8):
        pts = with_intensity(np.concatenate(static + [mover + np.float32([0.13 * f, 0, 0]), small]))
        perm = rng.permutation(len(pts))
        a.push_raw_cloud_and_pose(pts, IDENTITY_POSE)
        b.push_raw_cloud_and_pose(pts[perm], IDENTITY_POSE)
        ca, cb = a.tap("cluster_id"), b.tap("cluster_id")
        sets_a = {frozenset(np.flatnonzero(ca == k).tolist()) for k in range(a.counts()["K"])}
        sets_b = {frozenset(perm[np.flatnonzero(cb == k)].tolist()) for k in range(b.counts()["K"])}
        assert sets_a == sets_b and a.counts()["K"] == 5
        oa, ob = a.filter_cloud().copy(), b.filter_cloud().copy()
        assert sorted(map(bytes, oa)) == sorted(map(bytes, ob))
        assert a.counts()["NMO"] == b.counts()["NMO"]
    assert a.counts()["NMO"] == 1


def test_radius_tie_counter_counts_pairs_within_n_ulps_of_r2(oracle, tmp_path):
    """north_star: divergence at the exact clustering radius is to be counted. Pairs are placed along x at squared
    distances exactly r2, r2 -/+ 1 ulp, r2 -/+ 3 ulps (and far from each other): the counter must see exactly the ones
    inside the requested band."""
    cfg = write_cfg(tmp_path, ec_distance_threshold=0.25, min_cluster_size=1, **OPEN_CFG)
    r2 = np.float32(np.float64(np.float32(0.25)) ** 2)
    want = {0: r2}
    v = r2
    for u in range(1, 4):
        v = np.nextafter(v, np.float32(0))
        want[-u] = v
    v = r2
    for u in range(1, 4):
        v = np.nextafter(v, np.float32(10))
        want[u] = v
    pts = []
    found = {}
    for k, (u, d2) in enumerate(sorted(want.items())):
        base = np.float32(10.0 * k)  # pairs far apart from each other
        # search a float dx whose float square (the kernel's arithmetic with dy = dz = 0) equals d2 exactly
        dx = np.float32(np.sqrt(np.float64(d2)))
        for cand in (dx, np.nextafter(dx, np.float32(0)), np.nextafter(dx, np.float32(10))):
            if np.float32(cand * cand) == d2:
                found[u] = True
                pts += [[0.0, base, 0.0], [float(cand), base, 0.0]]
                break
    xyz = np.array(pts, np.float32)
    d2s = f32_sqdist_matrix(xyz)
    m = MovingObjectRemoval(cfg, 4, 3, binding=oracle)
    m.push_raw_cloud_and_pose(with_intensity(xyz), IDENTITY_POSE)
    r2f = np.float32(r2)
    for ulps in (0, 1, 2, 3):
        lo, hi = r2f, r2f
        for _ in range(ulps):
            lo, hi = np.nextafter(lo, np.float32(0)), np.nextafter(hi, np.float32(10))
        iu = np.triu_indices(len(xyz), 1)
        expect = int(np.sum((d2s[iu] >= lo) & (d2s[iu] <= hi)))
        assert m.radius_ties(ulps) == expect
    assert len(found) >= 3 and m.radius_ties(3) >= 3  # the construction did produce pairs inside the band
