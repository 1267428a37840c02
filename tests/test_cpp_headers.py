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
