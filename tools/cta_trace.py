"""Debug build only (libmor_b200_trace.so, -DMOR_CTA_TRACE): when did every CTA reach the barrier of every phase?
Per phase: the spread of the CTAs' arrival times relative to the phase's start (= the last arrival of the previous phase),
i.e. whether a phase is bound by one straggling CTA or by its common latency chain.
usage: MOR_PRODUCT_LIB=dynamicslamtool_b200/libmor_b200_trace.so python tools/cta_trace.py [frames...]"""
import ctypes as C, sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
b = load_product()
fn = b.lib.mor_debug_cta_trace
fn.argtypes = [C.c_void_p, C.POINTER(C.c_ulonglong), C.c_int]
want = [int(a) for a in sys.argv[1:]] or [20, 60, 104]
s = Synth(2, 2)
m = MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=b, max_points=s.max_points)
G = 148
for f in range(max(want) + 1):
    pts, pose = s.frame(f)
    m.push_raw_cloud_and_pose(pts, pose); m.filter_cloud()
    if f not in want: continue
    buf = (C.c_ulonglong * (32 * 256))()
    fn(m.h, buf, 32 * 256)
    t = np.frombuffer(buf, dtype=np.uint64).reshape(32, 256)[:, :G].astype(np.int64)
    names = list(m.phase_times().keys())
    c = m.counts()
    print(f"frame {f} NC {c['NC']} K {c['K']}")
    prev_end = None
    for ph, name in enumerate(names):
        arr = t[ph]
        if prev_end is None: prev_end = arr.min() - 1
        rel = (arr - prev_end) / 1e3
        order = np.argsort(rel)
        print(f"  {name:20s} first {rel.min():6.1f}  p50 {np.median(rel):6.1f}  p90 {np.percentile(rel, 90):6.1f}  last {rel.max():6.1f} us   slowest CTAs {order[-3:][::-1].tolist()}")
        if name == "ph_test":  # sub-steps (rows 16..18)
            for row, what in ((16, "start"), (17, "warp 0: light done"), (18, "light + heavy")):
                r2 = (t[row] - prev_end) / 1e3
                print(f"      .. {what:14s} p50 {np.median(r2):6.1f}  p90 {np.percentile(r2, 90):6.1f}  last {r2.max():6.1f}")
        if name.startswith("ph_stats"):  # the single-CTA tail (rows 19..22, slot 0): entry, centroids done, reciprocal NN done, end
            tt = [(t[row][0] - prev_end) / 1e3 for row in (19, 20, 21, 22)]
            print("      .. match: entry %.1f  centroids %.1f  nn %.1f  volume %.1f" % tuple(tt), " (nn searches done %.1f)" % ((t[23][0] - prev_end) / 1e3))
        if name.startswith("ph_moving"):  # chain tail (rows 24..27, slot 0): entry, flags + ring pushes, chains followed, pushCentroid done
            print("      .. chain: entry %.1f  flags+rings %.1f  follow %.1f  push %.1f" % tuple((t[row][0] - prev_end) / 1e3 for row in (24, 25, 26, 27)))
        prev_end = arr.max()
