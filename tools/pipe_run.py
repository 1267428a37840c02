"""A short pipelined device-resident run (mor_set_pipelining) for compute-sanitizer / ncu: frames pushed and filtered without
asking for anything, then one sync. usage: pipe_run.py [config] [scenario] [frames]"""
import ctypes as C, sys, zlib, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, Synth, load_product
cfg = sys.argv[1] if len(sys.argv) > 1 else 'config/MOR_config.txt'
scen = int(sys.argv[2]) if len(sys.argv) > 2 else 1
nfr = int(sys.argv[3]) if len(sys.argv) > 3 else 6
b = load_product()
s = Synth(scen, scen)
maxp = s.max_points
frames = [s.frame(f) for f in range(nfr)]
d = C.c_void_p()
assert b.device_alloc(0, nfr * maxp * 16, C.byref(d)) == 0
for f, (pts, _) in enumerate(frames):
    assert b.device_upload(0, C.c_void_p(d.value + f * maxp * 16), pts.ctypes.data_as(C.c_void_p), pts.nbytes) == 0
m = MovingObjectRemoval(cfg, 4, 3, binding=b, max_points=maxp)
m.set_pipelining(True)
for f, (pts, pose) in enumerate(frames):
    m.push_device(d.value + f * maxp * 16, len(pts), pose)
    m.filter_device(None, 0, want_count=False)
m.sync()
c = m.counts()
out = np.empty((c["NOUT"], 8), np.float32)
b.device_download(0, out.ctypes.data_as(C.c_void_p), C.c_void_p(m.output_device()), out.nbytes)
print(c, "crc %08x" % zlib.crc32(out.tobytes()))
