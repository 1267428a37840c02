"""The frame kernel's own phase timeline (us) on a few C2 frames."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
s = Synth(2, 2)
m = MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=load_product(), max_points=s.max_points)
for f in range(0, 108):
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.filter_cloud()
    if f in (5, 20, 40, 60, 104):
        t = m.phase_times()
        print(f, "NC", m.counts()["NC"], "total %.1f" % sum(t.values()), {k[3:]: round(v, 1) for k, v in t.items()})
