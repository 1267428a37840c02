"""Small driver for ncu captures and per-kernel timing experiments (not part of the product).
usage: prof_run.py [config] [scenario] [frames] [--profile] [--debug]   (--debug also fetches the VISUALIZE outputs)"""
import sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
args = [a for a in sys.argv[1:] if not a.startswith('--')]
cfg = args[0] if len(args) > 0 else 'config/MOR_config_hdl64.txt'
scen = int(args[1]) if len(args) > 1 else 2
nfr = int(args[2]) if len(args) > 2 else 12
s = Synth(scen, scen)
m = MovingObjectRemoval(cfg, 4, 3, binding=load_product(), max_points=s.max_points)
frames = [s.frame(f) for f in range(nfr)]
if '--profile' in sys.argv:
    m.set_kernel_profiling(True)
for pts, pose in frames:
    m.push_raw_cloud_and_pose(pts, pose)
    if '--debug' in sys.argv: m.cluster_collection()
    out = m.filter_cloud()
    if '--debug' in sys.argv: m.moving_markers()
print(m.counts())
if '--profile' in sys.argv:
    for k, (ms, n) in m.kernel_profile().items():
        if n: print(f"{k:24s} {1e3*ms/n:9.1f} us  x{n}")
