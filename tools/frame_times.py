"""Per-frame device time of push/filter (CUDA events) next to the frame's counts (not part of the product)."""
import sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
cfg, scen, nfr = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
s = Synth(scen, scen)
m = MovingObjectRemoval(cfg, 4, 3, binding=load_product(), max_points=s.max_points)
m.set_timing(True)
frames = [s.frame(f) for f in range(nfr)]
rows = []
for f, (pts, pose) in enumerate(frames):
    m.push_raw_cloud_and_pose(pts, pose); out = m.filter_cloud()
    p, fl = m.last_device_ms(); c = m.counts()
    rows.append((f, p * 1e3, fl * 1e3, c['NC'], c['K'], c['NK'], int(m.tap('cluster_size').max()) if c['K'] else 0))
for r in rows[:: max(1, nfr // 40)]:
    print("frame %4d push %8.1f us filter %6.1f us  NC %6d K %4d NK %6d maxcl %6d" % r)
a = np.array([r[1] for r in rows[5:]])
print("push us: mean %.1f p50 %.1f p99 %.1f max %.1f" % (a.mean(), np.percentile(a, 50), np.percentile(a, 99), a.max()))
