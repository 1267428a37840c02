import sys, numpy as np
sys.path.insert(0, '/root/repo')
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
cfg = sys.argv[1] if len(sys.argv) > 1 else 'config/MOR_config_hdl64.txt'
scen = int(sys.argv[2]) if len(sys.argv) > 2 else 2
nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 12
s = Synth(scen, scen)
m = MovingObjectRemoval(cfg, 4, 3, binding=load_product(), max_points=s.max_points)
for f in range(nfr):
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); out = m.filter_cloud()
print(m.counts())
