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


@pytest.mark.parametrize("scenario,cfg,frames,pipelining", [(1, "MOR_config.txt", 24, False), (2, "MOR_config_hdl64.txt", 40, False),
                                                            (1, "MOR_config.txt", 24, True), (2, "MOR_config_hdl64.txt", 40, True)])
def test_streamed_frames_equal_synchronous_frames_and_oracle(product, oracle, cfg_dir, scenario, cfg, frames, pipelining):
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
    m.set_pipelining(pipelining)  # the back half of frame f in one launch with the front half of frame f+1
    DEPTH = 4  # MOR_STREAM_DEPTH
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
    outs = [np.empty((s.max_points, 8), np.float32) for _ in range(5)]  # pageable memory works too (no overlap then)
    frames = [s.frame(f) for f in range(9)]
    for q in range(4):
        m.submit_frame(frames[q][0], frames[q][1], outs[q])
    with pytest.raises(MorError) as e:
        m.submit_frame(frames[4][0], frames[4][1], outs[4])
    assert e.value.status == 8  # MOR_STREAM_DEPTH = 4
    with pytest.raises(MorError) as e:
        m.push_raw_cloud_and_pose(*frames[4])
    assert e.value.status == 8  # synchronous calls only while nothing is in flight
    got = [m.collect_frame().copy() for _ in range(4)]
    # ... after which the synchronous calls continue the same sequence
    m.push_raw_cloud_and_pose(*frames[4]); got.append(m.filter_cloud().copy())
    m.submit_frame(frames[5][0], frames[5][1], outs[0]); got.append(m.collect_frame().copy())
    ref = MovingObjectRemoval(cfg_dir / "MOR_config.txt", 4, 3, binding=product, max_points=s.max_points)
    want = []
    for pts, pose in frames[:6]:
        ref.push_raw_cloud_and_pose(pts, pose); want.append(ref.filter_cloud().copy())
    for x, y in zip(got, want):
        assert np.array_equal(x, y)
    # an output buffer that is too small is reported at collection, with the needed count
    small = np.empty((16, 8), np.float32)
    m.submit_frame(frames[6][0], frames[6][1], small)
    with pytest.raises(MorError) as e:
        m.collect_frame()
    assert e.value.status == 6


@pytest.mark.parametrize("scenario,cfg,frames", [(1, "MOR_config.txt", 30), (2, "MOR_config_hdl64.txt", 36)])
def test_pipelined_device_resident_frames_equal_unpipelined(product, oracle, cfg_dir, scenario, cfg, frames):
    """mor_set_pipelining with the device-resident calls: frames enqueued without asking for anything run back half (f) beside
    front half (f+1) in one launch. Every few frames everything is compared with a handle that runs one kernel per frame and
    with the oracle (the comparison itself makes the pending back half run alone, so both ways of launching it are covered);
    the tracker state is a function of ALL frames so far, so agreement at the checkpoints covers the fused launches between."""
    from parity import ParityStats, compare_frame
    s = Synth(scenario, scenario)
    seq = [s.frame(f) for f in range(frames)]
    maxp = s.max_points
    pipe = MovingObjectRemoval(cfg_dir / cfg, 4, 3, binding=product, max_points=maxp)
    pipe.set_pipelining(True)
    orc = MovingObjectRemoval(cfg_dir / cfg, 4, 3, binding=oracle, max_points=maxp)
    d = C.c_void_p()
    assert product.device_alloc(0, len(seq) * maxp * 16, C.byref(d)) == 0
    for f, (pts, _) in enumerate(seq):
        assert product.device_upload(0, C.c_void_p(d.value + f * maxp * 16), pts.ctypes.data_as(C.c_void_p), pts.nbytes) == 0
    stats = ParityStats()
    checkpoints = {0, 1, 2, 5, 6, 13, frames - 2, frames - 1}
    for f, (pts, pose) in enumerate(seq):
        pipe.push_device(d.value + f * maxp * 16, len(pts), pose)
        pipe.filter_device(None, 0, want_count=False)
        orc.push_raw_cloud_and_pose(pts, pose)
        oo = orc.filter_cloud().copy()
        if f in checkpoints:
            c = pipe.counts()  # (flushes the pending back half)
            og = np.empty((c["NOUT"], 8), np.float32)
            if og.size:
                assert product.device_download(0, og.ctypes.data_as(C.c_void_p), C.c_void_p(pipe.output_device()), og.nbytes) == 0
            bad = compare_frame(pipe, orc, og, oo, stats)
            assert not bad, f"frame {f}: {bad}"
    assert stats.matches > 0
    product.device_free(0, d)
