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
