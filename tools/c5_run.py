"""C5: many independent C2-shaped sequences, seeds 1000, 1001, ..., a fixed number of frames each,
spread over the ranks of a torchrun job (sequence s -> rank s mod world, no collective on the data path).

    python tools/c5_run.py --sequences 256 --frames 32 --batch 16
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
        tools/c5_run.py --sequences 1024 --frames 32 --batch 16

Every rank takes its sequences `batch` at a time: the frames of the batch are generated on the host and staged in
HBM (untimed - the generator is not the subject), the handles are reset, and the batch is advanced `frames` times
with mor_batch_step_device, timed with CUDA events on the launching stream. The figure is all frames of all ranks
divided by the largest per-rank sum of timed milliseconds. Parity of the batched path is the
business of tests/test_gpu_edge_cases.py (batched stepping against the oracle); this tool only measures.
"""
import argparse
import ctypes as C
import json
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path


ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dynamicslamtool_b200 import MovingObjectRemoval, SequenceBatch, Synth, load_product  # noqa: E402
from dynamicslamtool_b200.replicas import combine, seed_for_sequence, sequences_for_rank  # noqa: E402

CFG = ROOT / "config" / "MOR_config_hdl64.txt"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=256)
    ap.add_argument("--frames", type=int, default=32)
    ap.add_argument("--batch", type=int, default=37)
    ap.add_argument("--base-seed", type=int, default=1000)
    ap.add_argument("--threads", type=int, default=0)
    args = ap.parse_args()

    rank, world, local_rank = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    import torch
    torch.cuda.set_device(local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    b = load_product()
    mine = sequences_for_rank(args.sequences, rank, world)
    S, T = args.batch, args.frames
    maxp = Synth(2, args.base_seed).max_points
    frame_bytes = maxp * 16
    threads = args.threads or max(1, (os.cpu_count() or 8) // max(1, min(world, 8)))

    d_in = C.c_void_p()
    assert b.device_alloc(local_rank, S * T * frame_bytes, C.byref(d_in)) == 0, "staging allocation failed"
    d_outs = []
    for _ in range(S):
        p = C.c_void_p()
        assert b.device_alloc(local_rank, maxp * 32, C.byref(p)) == 0
        d_outs.append(p.value)
    hs = [MovingObjectRemoval(CFG, 4, 3, device=local_rank, binding=b, max_points=maxp) for _ in range(S)]

    def generate(seq):
        syn = Synth(2, seed_for_sequence(args.base_seed, seq))
        syn.threads = 1  # the pool below already runs one sequence per host thread
        return [syn.frame(f) for f in range(T)]

    timed_ms, frames_done, launches, points_in, points_out, gen_s = 0.0, 0, 0, 0, 0, 0.0
    for b0 in range(0, len(mine), S):
        seqs = mine[b0:b0 + S]
        t0 = time.time()
        with ThreadPoolExecutor(threads) as ex:
            data = list(ex.map(generate, seqs))
        gen_s += time.time() - t0
        for si, frames in enumerate(data):
            for f, (pts, _) in enumerate(frames):
                assert b.device_upload(local_rank, C.c_void_p(d_in.value + (si * T + f) * frame_bytes), pts.ctypes.data_as(C.c_void_p), pts.nbytes) == 0
        group = hs[:len(seqs)]
        for h in group:
            h.reset()
        batch = SequenceBatch(group)
        lead = group[0]
        l0 = lead.launch_count()
        lead.event_record(0)
        for f in range(T):
            batch.step_device([d_in.value + (si * T + f) * frame_bytes for si in range(len(seqs))], [int(data[si][f][0].shape[0]) for si in range(len(seqs))],
                              [data[si][f][1] for si in range(len(seqs))], d_outs[:len(seqs)])
        lead.event_record(1)
        lead.sync()
        timed_ms += lead.event_elapsed_ms(0, 1)
        launches += lead.launch_count() - l0
        frames_done += len(seqs) * T
        for si, h in enumerate(group):
            h.sync()
            c = h.counts()
            if c["ERRFLAGS"]:
                raise RuntimeError(f"sequence {seqs[si]}: device capacity flags {c['ERRFLAGS']}")
            points_out += c["NOUT"]
            points_in += sum(int(fr[0].shape[0]) for fr in data[si])

    total_frames, worst_ms = combine(frames_done, timed_ms, torch.device("cuda", local_rank))
    total_launches, _ = combine(launches, 0.0, torch.device("cuda", local_rank))
    total_in, _ = combine(points_in, 0.0, torch.device("cuda", local_rank))
    if rank == 0:
        print(json.dumps({
            "workload": f"C5: {args.sequences} independent C2-shaped sequences (seeds {args.base_seed}..{args.base_seed + args.sequences - 1}), {T} frames each, "
                        f"{S} sequences per set of launches, config MOR_config_hdl64.txt",
            "n_gpus": world, "frames": total_frames, "value": total_frames / (worst_ms * 1e-3), "unit": "frames/s", "timed_ms_max_over_ranks": worst_ms,
            "points_per_frame": total_in / max(1, total_frames), "gpu_launches": total_launches, "data": "synthetic, staged in HBM before the timed region",
            "output_points_last_frames": points_out, "generator_seconds_rank0": gen_s}))
    for h in hs:
        h.close()
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
