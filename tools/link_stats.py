import sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
s = Synth(2, 2)
m = MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=load_product(), max_points=s.max_points)
m.set_timing(True)
for f in range(0, 180):
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.sync()
    d = m.tap('debug_scratch')[4:]
    p, _ = 0, 0
    m.filter_cloud()
    if f % 5 == 0:
        c = m.counts()
        print(f"f{f:4d} NC {c['NC']:6d} tested {d[0]:9d} scans {d[1]:8d} maxscan {d[2]:6d} fullmiss {d[3]:7d} fullmiss_pts {d[4]:9d} hits {d[5]:7d} boxpruned {d[6]:7d}")
