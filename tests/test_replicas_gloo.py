"""world_size-2 gloo test of the replica bookkeeping (the only multi-process logic on the path)."""
import ctypes as C
import os
import sys
from pathlib import Path

import numpy as np
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from dynamicslamtool_b200 import MorBinding, MovingObjectRemoval, Synth
    from dynamicslamtool_b200.replicas import combine, seed_for_sequence, sequences_for_rank
    orc = MorBinding(C.CDLL(str(ROOT / "oracle" / "libmor_oracle.so")), "oracle_")
    mine = sequences_for_rank(3, rank, world)
    frames, checks = 0, []
    for s in mine:  # each owned sequence: 2 frames of C1 through the oracle (CPU stand-in for the device path)
        m = MovingObjectRemoval(ROOT / "config" / "MOR_config.txt", 4, 3, binding=orc)
        syn = Synth(1, seed_for_sequence(100, s))
        for f in range(2):
            pts, pose = syn.frame(f)
            m.push_raw_cloud_and_pose(pts, pose)
            out = m.filter_cloud()
            frames += 1
            checks.append((s, f, int(out.shape[0])))
    total, worst = combine(frames, 10.0 * (rank + 1))
    q.put((rank, mine, total, worst, checks))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_replicas(built):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, s0, tot0, w0, c0), (r1, s1, tot1, w1, c1) = res
    assert s0 == [0, 2] and s1 == [1]                 # sequence s -> rank s mod world, no overlap, full cover
    assert tot0 == tot1 == 6                            # 3 sequences x 2 frames, summed over ranks
    assert w0 == w1 == 20.0                             # max over ranks
    # sequences are independent: rank 1's result for sequence 1 does not depend on who else ran
    assert all(n > 0 for _, _, n in c0 + c1)
