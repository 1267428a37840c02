"""S sequences of C2 stepped with mor_batch_step_device (for ncu launch lists of the batched kernels)."""
import ctypes as C, sys, numpy as np
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from dynamicslamtool_b200 import MovingObjectRemoval, SequenceBatch, Synth, load_product
S = int(sys.argv[1]) if len(sys.argv) > 1 else 16
T = int(sys.argv[2]) if len(sys.argv) > 2 else 12
b = load_product()
s = Synth(2, 2)
maxp = s.max_points
F = 24
frames = [s.frame(f) for f in range(F)]
d_frames = []
for pts, _ in frames:
    p = C.c_void_p(); assert b.device_alloc(0, maxp * 16, C.byref(p)) == 0
    b.device_upload(0, p, pts.ctypes.data_as(C.c_void_p), pts.nbytes); d_frames.append(p)
hs = [MovingObjectRemoval('config/MOR_config_hdl64.txt', 4, 3, binding=b, max_points=maxp) for _ in range(S)]
outs = []
for _ in range(S):
    p = C.c_void_p(); assert b.device_alloc(0, maxp * 32, C.byref(p)) == 0; outs.append(p.value)
batch = SequenceBatch(hs)
for t in range(T):
    fs = [(si + t) % F for si in range(S)]   # consecutive frames per sequence, different phase per sequence
    batch.step_device([d_frames[f].value for f in fs], [frames[f][0].shape[0] for f in fs], [frames[f][1] for f in fs], outs)
hs[0].sync()
print(hs[0].counts())
