"""mor_submit_frame / mor_collect_frame (pipelined streaming through the C ABI): the frames delivered one call late must
be byte-identical to what pushRawCloudAndPose + filterCloud deliver frame by frame, and to the oracle's."""
import ctypes as C

import numpy as np
import pytest

from dynamicslamtool_b200 import MorError, MovingObjectRemoval, Synth
from helpers import crc

pytestmark = pytest.mark.gpu


def _pinned(b, shape, dtype=np.float32):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    p = C.c_void_p()
    assert b.alloc_pinned(nbytes, C.byref(p)) == 0
    buf = (C.c_uint8 * nbytes).from_address(p.value)
    return np.frombuffer(buf, dtype=dtype).reshape(shape), p


@pytest.mark.parametrize("scenario,cfg,frames", [(1, "MOR_config.txt", 24), (2, "MOR_config_hdl64.txt", 40)])
def test_streamed_frames_equal_synchronous_frames_and_oracle(product, oracle, cfg_dir, scenario, cfg, frames):
    s = Synth(scenario, scenario)
    seq = [s.frame(f) for f in range(frames)]
    ref = MovingObjectRemoval(cfg_dir / cfg, 4, 3, binding=product, max_points=s.max_points)
    orc = MovingObjectRemoval(cfg_dir / cfg, 4, 3, binding=oracle, max_points=s.max_points)
    want, want_orc = [], []
    for pts, pose in seq:
        ref.push_raw_cloud_and_pose(pts, pose)
        want.append(crc(ref.filter_cloud()))
        orc.push_raw_cloud_and_pose(pts, pose)
        want_orc.append(crc(orc.filter_cloud()))
    assert want == want_orc
    m = MovingObjectRemoval(cfg_dir / cfg, 4, 3, binding=product, max_points=s.max_points)
    DEPTH = 3  # MOR_STREAM_DEPTH
    ins, outs, keep = [], [], []
    for _ in range(DEPTH):
        a, pa = _pinned(product, (s.max_points, 4)); o, po = _pinned(product, (s.max_points, 8))
        ins.append(a); outs.append(o); keep += [pa, po]
    got = []
    for f, (pts, pose) in enumerate(seq):
        slot = f % DEPTH
        ins[slot][: len(pts)] = pts
        m.submit_frame(ins[slot][: len(pts)], pose, outs[slot])
        assert m.frames_in_flight() == min(f + 1, DEPTH)
        if f >= DEPTH - 1:
            got.append(crc(m.collect_frame()))
    while m.frames_in_flight():
        got.append(crc(m.collect_frame()))
    assert m.frames_in_flight() == 0
    assert got == want
    # the tracker state at the end equals the synchronous handle's
    assert np.array_equal(m.tap("mo_conf"), ref.tap("mo_conf")) and np.array_equal(m.tap("mo_centroids"), ref.tap("mo_centroids"))
    for p in keep:
        product.free_pinned(p)


def test_streaming_protocol_errors_and_mixing(product, cfg_dir):
    s = Synth(1, 1)
    m = MovingObjectRemoval(cfg_dir / "MOR_config.txt", 4, 3, binding=product, max_points=s.max_points)
    with pytest.raises(MorError) as e:
        m.collect_frame()
    assert e.value.status == 8  # nothing in flight
    outs = [np.empty((s.max_points, 8), np.float32) for _ in range(4)]  # pageable memory works too (no overlap then)
    frames = [s.frame(f) for f in range(8)]
    for q in range(3):
        m.submit_frame(frames[q][0], frames[q][1], outs[q])
    with pytest.raises(MorError) as e:
        m.submit_frame(frames[3][0], frames[3][1], outs[3])
    assert e.value.status == 8  # MOR_STREAM_DEPTH = 3
    with pytest.raises(MorError) as e:
        m.push_raw_cloud_and_pose(*frames[3])
    assert e.value.status == 8  # synchronous calls only while nothing is in flight
    got = [m.collect_frame().copy() for _ in range(3)]
    # ... after which the synchronous calls continue the same sequence
    m.push_raw_cloud_and_pose(*frames[3]); got.append(m.filter_cloud().copy())
    m.submit_frame(frames[4][0], frames[4][1], outs[0]); got.append(m.collect_frame().copy())
    ref = MovingObjectRemoval(cfg_dir / "MOR_config.txt", 4, 3, binding=product, max_points=s.max_points)
    want = []
    for pts, pose in frames[:5]:
        ref.push_raw_cloud_and_pose(pts, pose); want.append(ref.filter_cloud().copy())
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
    # an output buffer that is too small is reported at collection, with the needed count
    small = np.empty((16, 8), np.float32)
    m.submit_frame(frames[5][0], frames[5][1], small)
    with pytest.raises(MorError) as e:
        m.collect_frame()
    assert e.value.status == 6
