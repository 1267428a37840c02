"""Debug build only (make EXTRA=-DMOR_LINK_STATS): union-find and heavy-pair counters of the link phases, per frame."""
import ctypes as C, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
b = load_product()
fn = b.lib.mor_debug_link_stats
fn.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
s = Synth(2, 2)
m = MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=b, max_points=s.max_points)
names = ["cells", "rounds", "max_cell_cyc", "cyc_findA", "cyc_check", "cyc_unions", "union_rounds", "real_unions", "hit_neighbours", "heavy_same_root", "heavy_hit", "heavy_miss",
         "cyc_probe", "cyc_box_append", "cyc_list_test_union", "cyc_union_only"]
for f in list(range(0, 2)) + [20, 60]:
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.filter_cloud()
    out = (C.c_ulonglong * 16)()
    fn(out, 1)
    c = m.counts()
    print(f, "NC", c["NC"], {n: int(out[i]) for i, n in enumerate(names)}, {k: round(v, 1) for k, v in m.phase_times().items() if "link" in k})
