"""CPU: the C ABI header compiles as plain C and the class header as C++ without ROS/PCL."""
import subprocess

from helpers import ROOT


def test_c_header_is_plain_c(tmp_path):
    src = tmp_path / "t.c"
    src.write_text('#include "mor_b200.h"\nint main(void){ mor_limits l = {0}; (void)l; return MOR_TAP__COUNT > 20 ? 0 : 1; }\n')
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_class_header_compiles_without_ros(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text('#include "MOR/MovingObjectRemoval.h"\nint f(MovingObjectRemoval& m, pcl::PCLPointCloud2& c, geometry_msgs::Pose p){ m.pushRawCloudAndPose(c, p); return m.filterCloud(c, "/filtered") ? (int)m.output.width : -1; }\n')
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-I", str(ROOT / "include"), "-c", str(src), "-o", str(tmp_path / "t.o")])


def test_replay_front_end_pose_conversion(built):
    """CPU: the 3x4 matrix -> pose7 conversion of the harness's replay front end (harness/replay_io.h) against scipy,
    on every branch of the largest-component rule (angles near 0 and near 180 degrees about each axis)."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    exe = ROOT / "harness" / "mov_harness"
    subprocess.check_call(["make", "-C", str(ROOT / "harness")], stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(5)
    rots = [Rotation.random(random_state=int(rng.integers(1 << 30))) for _ in range(6)]
    rots += [Rotation.from_rotvec(np.array(ax) * ang) for ax in ([1, 0, 0], [0, 1, 0], [0, 0, 1]) for ang in (1e-9, np.pi - 1e-6, np.pi)]
    for r in rots:
        m = np.concatenate([r.as_matrix(), rng.normal(size=(3, 1)) * 10], axis=1)
        out = subprocess.run([str(exe), "--pose-of"] + [repr(float(v)) for v in m.reshape(-1)], capture_output=True, text=True, timeout=60)
        assert out.returncode == 0, out.stderr
        p = np.array([float(v) for v in out.stdout.split()])
        assert np.array_equal(p[:3], m[:, 3])
        q = r.as_quat()
        assert min(np.abs(p[3:] - q).max(), np.abs(p[3:] + q).max()) < 1e-7
        assert abs(np.linalg.norm(p[3:]) - 1) < 1e-12


def test_replay_front_end_reads_recorded_directories(built, tmp_path):
    """CPU: the recorded-data front end (harness/replay_io.h) through `mov_harness --inspect`: .bin clouds in lexical
    order (also under velodyne/), 7- / 8- / 12-column pose files, comments, calib.txt Tr, clouds without a pose
    dropped, malformed inputs rejected."""
    import numpy as np
    from scipy.spatial.transform import Rotation
    exe = ROOT / "harness" / "mov_harness"
    subprocess.check_call(["make", "-C", str(ROOT / "harness")], stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(9)

    def inspect(d):
        return subprocess.run([str(exe), "--inspect", str(d)], capture_output=True, text=True, timeout=60)

    def rows(res):
        return [l.split() for l in res.stdout.splitlines() if l.startswith("frame ")]

    sizes = [5, 11, 3, 8]
    poses = [np.concatenate([rng.normal(size=3), Rotation.random(random_state=i).as_quat()]) for i in range(4)]
    # 7 columns, clouds directly in the directory, one cloud more than poses
    d = tmp_path / "seven"
    d.mkdir()
    for i, n in enumerate(sizes + [2]):
        rng.normal(size=(n, 4)).astype(np.float32).tofile(d / f"{i:06d}.bin")
    (d / "poses.txt").write_text("# comment\n" + "\n".join(" ".join(repr(float(v)) for v in p) for p in poses) + "\n")
    res = inspect(d)
    assert res.returncode == 0, res.stderr
    assert res.stdout.splitlines()[0] == "frames 4 max_points 11"
    got = rows(res)
    assert [int(r[3]) for r in got] == sizes
    assert np.array_equal(np.array([[float(v) for v in r[5:12]] for r in got]), np.array(poses))
    # 8 columns (TUM), clouds under velodyne/
    d = tmp_path / "tum"
    (d / "velodyne").mkdir(parents=True)
    for i, n in enumerate(sizes):
        rng.normal(size=(n, 4)).astype(np.float32).tofile(d / "velodyne" / f"{i:06d}.bin")
    (d / "poses.txt").write_text("\n".join(" ".join([repr(0.1 * i)] + [repr(float(v)) for v in p]) for i, p in enumerate(poses)) + "\n")
    got = rows(inspect(d))
    assert np.array_equal(np.array([[float(v) for v in r[5:12]] for r in got]), np.array(poses))
    # 12 columns + calib: sensor pose = Tr^-1 P Tr
    d = tmp_path / "kitti"
    d.mkdir()
    tr = np.eye(4)
    tr[:3, :3] = Rotation.from_euler("xyz", [-90, 0, -90], degrees=True).as_matrix()
    tr[:3, 3] = [0.1, -0.2, 0.3]
    lines = []
    for i, (n, p) in enumerate(zip(sizes, poses)):
        rng.normal(size=(n, 4)).astype(np.float32).tofile(d / f"{i:06d}.bin")
        m = np.eye(4)
        m[:3, :3] = Rotation.from_quat(p[3:]).as_matrix()
        m[:3, 3] = p[:3]
        lines.append(" ".join(repr(float(v)) for v in (tr @ m @ np.linalg.inv(tr))[:3].reshape(-1)))
    (d / "poses.txt").write_text("\n".join(lines) + "\n")
    (d / "calib.txt").write_text("P0: 1 0 0 0 0 1 0 0 0 0 1 0\nTr: " + " ".join(repr(float(v)) for v in tr[:3].reshape(-1)) + "\n")
    got = rows(inspect(d))
    for r, p in zip(got, poses):
        q = np.array([float(v) for v in r[5:12]])
        assert np.allclose(q[:3], p[:3], atol=1e-9)
        assert min(np.abs(q[3:] - p[3:]).max(), np.abs(q[3:] + p[3:]).max()) < 1e-9
    # malformed: a pose line with 5 numbers; a cloud whose size is not a multiple of 16 bytes; an empty directory
    (d / "poses.txt").write_text("1 2 3 4 5\n")
    assert inspect(d).returncode != 0
    d2 = tmp_path / "badbin"
    d2.mkdir()
    (d2 / "000000.bin").write_bytes(b"\0" * 20)
    (d2 / "poses.txt").write_text("0 0 0 0 0 0 1\n")
    assert inspect(d2).returncode != 0
    d3 = tmp_path / "empty"
    d3.mkdir()
    assert inspect(d3).returncode != 0


def test_ros_branch_and_node_type_check_against_stub_headers():
    """CPU: the MOR_WITH_ROS branch of the class header, the class implementation and the ROS node ros/mov_e.cpp
    (reference: src/external_sync_test.cpp:7-41 + the VISUALIZE publishers of cpp:553-558, :640-642) compile against
    minimal stand-ins of the ROS / PCL headers (ros/stubs/): the branch is at least type-checked offline."""
    inc = ["-I", str(ROOT / "ros" / "stubs"), "-I", str(ROOT / "include")]
    for src in (ROOT / "ros" / "mov_e.cpp", ROOT / "dynamicslamtool_b200" / "csrc" / "MovingObjectRemoval.cpp"):
        subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only", "-DMOR_WITH_ROS"] + inc + [str(src)])
    # and the node is an empty program without ROS
    subprocess.check_call(["g++", "-std=c++14", "-Wall", "-Werror", "-fsyntax-only", "-I", str(ROOT / "include"), str(ROOT / "ros" / "mov_e.cpp")])


def _pairs(exe, tmp_path, ta, tb, queue=10):
    import numpy as np
    fa, fb = tmp_path / "a.txt", tmp_path / "b.txt"
    fa.write_text("\n".join(repr(float(t)) for t in ta) + "\n")
    fb.write_text("\n".join(repr(float(t)) for t in tb) + "\n")
    out = subprocess.run([str(exe), "--pair", str(fa), str(fb), str(queue)], capture_output=True, text=True, timeout=60)
    assert out.returncode == 0, out.stderr
    return np.array([[int(v) for v in l.split()[1:]] for l in out.stdout.splitlines() if l.startswith("pair ")], int).reshape(-1, 2)


def test_approximate_time_pairing(built, tmp_path):
    """CPU: the harness's ApproximateTime(10) pairing of independently stamped cloud / odometry streams (the policy the
    reference node uses, src/external_sync_test.cpp:31-35): identical stamps pair one to one; a 10 Hz stream against a
    jittered 100 Hz stream pairs every cloud with its nearest odometry sample; pairs never reuse or reorder messages;
    a stalled topic overflows the queue of the other without breaking any of that."""
    import numpy as np
    exe = ROOT / "harness" / "mov_harness"
    subprocess.check_call(["make", "-C", str(ROOT / "harness")], stdout=subprocess.DEVNULL)
    rng = np.random.default_rng(3)
    t = np.arange(50) * 0.1
    p = _pairs(exe, tmp_path, t, t)
    assert np.array_equal(p, np.stack([np.arange(50), np.arange(50)], 1))
    # constant offset smaller than half a period: still one to one
    p = _pairs(exe, tmp_path, t, t + 0.03)
    assert np.array_equal(p[:, 0], p[:, 1]) and len(p) >= 48
    # 10 Hz clouds, ~100 Hz odometry with jitter
    ta = 0.05 + np.arange(60) * 0.1 + rng.uniform(-0.002, 0.002, 60)
    tb = np.sort(np.arange(700) * 0.01 + rng.uniform(-0.003, 0.003, 700))
    p = _pairs(exe, tmp_path, ta, tb)
    assert len(p) >= 57
    assert np.all(np.diff(p[:, 0]) > 0) and np.all(np.diff(p[:, 1]) > 0)          # never reused, never reordered
    nearest = np.abs(ta[p[:, 0], None] - tb[None, :]).argmin(1)
    assert np.array_equal(p[:, 1], nearest)                                       # each cloud with its nearest odometry sample
    # odometry stalls for 3 s, the cloud queue (10) overflows, then both resume
    tb2 = np.concatenate([tb[tb < 2.0], tb[tb > 5.0]])
    p = _pairs(exe, tmp_path, ta, tb2)
    assert np.all(np.diff(p[:, 0]) > 0) and np.all(np.diff(p[:, 1]) > 0)
    late = p[ta[p[:, 0]] > 5.2]
    assert len(late) >= 5 and np.all(np.abs(ta[late[:, 0]] - tb2[late[:, 1]]) < 0.008)
    assert not np.any((ta[p[:, 0]] > 2.2) & (ta[p[:, 0]] < 3.9))                  # clouds of the stall that fell out of the queue are never paired


def test_unsynchronised_recording_is_paired_before_replay(built, tmp_path):
    """CPU: --inspect on a directory with times.txt + odometry.txt (no poses.txt): the frames fed to the class are the
    ApproximateTime pairs, each cloud with the odometry pose nearest in time."""
    import numpy as np
    exe = ROOT / "harness" / "mov_harness"
    rng = np.random.default_rng(4)
    d = tmp_path / "rec"
    d.mkdir()
    n = 12
    for i in range(n):
        rng.normal(size=(4 + i, 4)).astype(np.float32).tofile(d / f"{i:06d}.bin")
    tc = 100.0 + np.arange(n) * 0.1
    to = 99.95 + np.arange(140) * 0.01
    (d / "times.txt").write_text("\n".join(repr(float(v)) for v in tc) + "\n")
    odom = np.concatenate([to[:, None], rng.normal(size=(140, 3)), np.tile([0, 0, 0, 1.0], (140, 1))], 1)
    (d / "odometry.txt").write_text("\n".join(" ".join(repr(float(v)) for v in r) for r in odom) + "\n")
    res = subprocess.run([str(exe), "--inspect", str(d)], capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr
    rows = [l.split() for l in res.stdout.splitlines() if l.startswith("frame ")]
    assert len(rows) >= n - 1
    for r in rows:
        k = int(r[3]) - 4                                    # which cloud (by its size)
        j = np.abs(to - tc[k]).argmin()
        assert np.allclose([float(v) for v in r[5:8]], odom[j, 1:4])
