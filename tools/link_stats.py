import sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
s = Synth(2, 2)
m = MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=load_product(), max_points=s.max_points)
for f in range(0, 125):
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.sync()
    d = m.tap('debug_scratch')
    d = d[-8:]
    m.filter_cloud()
    if f % 8 == 0 or f in (103,105,107):
        c = m.counts()
        print(f"f{f:4d} NC {c['NC']:6d} thr max {d[0]:8d} cyc mean {d[1]/max(d[2],1):8.0f} n {d[2]:7d} >20k {d[6]:6d} >40k {d[7]:6d} | union max {d[3]:8d} mean {d[4]/max(d[5],1):8.0f} n {d[5]:6d}")
