"""Long parity soak: N frames of a scenario through the CUDA path and the oracle, all taps compared per frame."""
import ctypes as C, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from dynamicslamtool_b200 import MorBinding, MovingObjectRemoval, Synth, load_product
from parity import ParityStats, compare_frame
cfg, scen, n = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
method = int(sys.argv[4]) if len(sys.argv) > 4 else 2
if method != 2:
    text = Path(cfg).read_text().replace("method_choice:2", f"method_choice:{method}")
    cfg = "/tmp/soak_cfg.txt"; Path(cfg).write_text(text)
orc = MorBinding(C.CDLL(str(ROOT / "oracle" / "libmor_oracle.so")), "oracle_")
s = Synth(scen, scen)
g = MovingObjectRemoval(cfg, 4, 3, binding=load_product(), max_points=s.max_points)
o = MovingObjectRemoval(cfg, 4, 3, binding=orc)
st = ParityStats(); t0 = time.time()
for f in range(n):
    pts, pose = s.frame(f)
    g.push_raw_cloud_and_pose(pts, pose); o.push_raw_cloud_and_pose(pts, pose)
    og, oo = g.filter_cloud().copy(), o.filter_cloud().copy()
    bad = compare_frame(g, o, og, oo, st)
    if bad:
        print("DIVERGENCE at frame", f, bad); sys.exit(1)
print("soak ok", cfg, "scenario", scen, "frames", n, st.as_dict(), "%.0fs" % (time.time() - t0))
