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
names = ["cells", "-", "max_cell_cyc", "hv_rootcheck_cyc", "hv_rootcheck_max", "hv_test_cyc", "hv_union_cyc", "heavy_pairs", "hv_pair_max_cyc", "heavy_same_root", "heavy_hit", "heavy_miss",
         "cyc_probe", "cyc_box_append", "cyc_list_test_union", "cyc_union_only"]
for f in list(range(0, 3)) + [20, 60, 104]:
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.filter_cloud()
    out = (C.c_ulonglong * 16)()
    fn(out, 1)
    c = m.counts()
    print(f, "NC", c["NC"], {n: int(out[i]) for i, n in enumerate(names)}, {k: round(v, 1) for k, v in m.phase_times().items() if "link" in k})
