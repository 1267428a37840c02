"""Per-config summary: frames/s through the host C ABI (pinned buffers, H2D+D2H inside) and
device-resident, p50/p99 latency, CPU oracle on one core over a bounded sample, algorithmic-bytes roofline
fraction. Not the bench contract (bench.py is); a convenience table for C1..C4."""
import ctypes as C
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from bench import algorithmic_bytes, measured_peak_gbs, pinned_array  # noqa: E402
from dynamicslamtool_b200 import MorBinding, MovingObjectRemoval, Synth, load_product  # noqa: E402

CONFIGS = [("C1", 1, "MOR_config.txt", 100), ("C2", 2, "MOR_config_hdl64.txt", 200), ("C3", 3, "MOR_config_os128.txt", 100),
           ("C4", 4, "MOR_config_terrain.txt", 100)]


def main():
    b = load_product()
    orc = MorBinding(C.CDLL(str(ROOT / "oracle" / "libmor_oracle.so")), "oracle_")
    peak, _ = measured_peak_gbs()
    rows = []
    for name, scen, cfgname, frames in CONFIGS:
        cfg = ROOT / "config" / cfgname
        s = Synth(scen, scen)
        maxp = s.max_points
        pts, _p = pinned_array(b, (frames, maxp, 4), np.float32)
        out, _o = pinned_array(b, (maxp, 8), np.float32)
        npts, poses = np.zeros(frames, np.int64), np.zeros((frames, 7))
        for f in range(frames):
            p, poses[f] = s.frame(f, out=pts[f])
            npts[f] = p.shape[0]
        W = 10
        m = MovingObjectRemoval(cfg, 4, 3, binding=b, max_points=maxp)
        lat, alg = [], []
        for f in range(frames):
            t0 = time.perf_counter()
            m.push_raw_cloud_and_pose(pts[f, : npts[f]], poses[f])
            m.filter_cloud(out)
            dt = time.perf_counter() - t0
            if f >= W:
                lat.append(dt)
                alg.append(algorithmic_bytes(m.counts()))
        m.close()
        lat = np.array(lat)
        d_frames, d_out = C.c_void_p(), C.c_void_p()
        b.device_alloc(0, frames * maxp * 16, C.byref(d_frames)); b.device_alloc(0, maxp * 32, C.byref(d_out))
        b.device_upload(0, d_frames, pts.ctypes.data_as(C.c_void_p), frames * maxp * 16)
        m = MovingObjectRemoval(cfg, 4, 3, binding=b, max_points=maxp)
        for f in range(W):
            m.push_device(d_frames.value + f * maxp * 16, int(npts[f]), poses[f]); m.filter_device(d_out.value, maxp, want_count=False)
        m.sync(); m.event_record(0)
        for f in range(W, frames):
            m.push_device(d_frames.value + f * maxp * 16, int(npts[f]), poses[f]); m.filter_device(d_out.value, maxp, want_count=False)
        m.event_record(1)
        dev_ms = m.event_elapsed_ms(0, 1) / (frames - W)
        m.close(); b.device_free(0, d_frames); b.device_free(0, d_out)
        mo = MovingObjectRemoval(cfg, 4, 3, binding=orc)
        t_used, nf = 0.0, 0
        oo = np.empty((maxp, 8), np.float32)
        for f in range(frames):
            t0 = time.perf_counter(); mo.push_raw_cloud_and_pose(pts[f, : npts[f]], poses[f]); mo.filter_cloud(oo); t_used += time.perf_counter() - t0; nf += 1
            if t_used > 12:
                break
        row = dict(config=name, cfg=cfgname, points=int(npts.mean()), cpu_fps=nf / t_used, cpu_frames=nf, e2e_fps=1.0 / lat.mean(), p50_ms=float(np.percentile(lat, 50) * 1e3),
                   p99_ms=float(np.percentile(lat, 99) * 1e3), dev_fps=1e3 / dev_ms, alg_mb=float(np.mean(alg)) / 1e6, frac=float(np.mean(alg)) / (dev_ms * 1e-3) / 1e9 / peak)
        rows.append(row)
        print(json.dumps(row), flush=True)
    print("| config | points/frame | CPU oracle 1-core fps | B200 e2e fps | p50 ms | p99 ms | B200 device-resident fps | alg. MB/frame | HBM roofline frac |")
    print("|---|---|---|---|---|---|---|---|---|")
    for r in rows:
        print(f"| {r['config']} ({r['cfg']}) | {r['points']} | {r['cpu_fps']:.2f} ({r['cpu_frames']} frames) | {r['e2e_fps']:.0f} | {r['p50_ms']:.3f} | {r['p99_ms']:.3f} | {r['dev_fps']:.0f} | {r['alg_mb']:.1f} | {r['frac']:.4f} |")


if __name__ == "__main__":
    main()
