"""Row-band split of the CBLSM path (ComputeAD -> CrossAggregator::Aggregate(4)) over the ranks of a torchrun job.

    python -m torch.distributed.run --nproc-per-node N ... scripts/bandsplit_check.py --shape HxWxD [--verify rank0|none]

Every rank builds the same seeded image pair, computes the AD cost of its own band (+34-row halos) and aggregates it with
one neighbour halo exchange per iteration (stereo_match_traditional_b200/bandsplit.py).  --verify rank0: rank 0 also
aggregates the WHOLE frame on its own GPU and compares its band; --verify all: every band is gathered and compared."""
import argparse
import json
import os
import sys

sys.path.insert(0, ".")
import numpy as np
import torch
import torch.distributed as dist

from stereo_match_traditional_b200.api import SUM_FAST, Engine
from stereo_match_traditional_b200.bandsplit import aggregate_cross_band, make_band
from stereo_match_traditional_b200.synth import make_pair


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="1080x1920x64")
    ap.add_argument("--verify", default="all", choices=["all", "rank0", "none"])
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    H, W, D = (int(x) for x in a.shape.split("x"))
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # a modest tile repeated over the frame keeps host generation cheap at 8K while every rank gets identical data
    th, tw = min(H, 540), min(W, 960)
    p = make_pair(th, tw, min(D, 64), 1005, "scene")
    reps = (-(-H // th), -(-W // tw))
    bgr = torch.as_tensor(np.tile(p["bgr_left"], reps + (1,))[:H, :W].copy(), device=dev)
    gl = torch.as_tensor(np.tile(p["gray_left"], reps)[:H, :W].copy(), device=dev)
    gr = torch.as_tensor(np.tile(p["gray_right"], reps)[:H, :W].copy(), device=dev)
    eng = Engine(local)
    band = make_band(rank, world, H)
    # CBLSM ComputeAD on the band's rows (row-local)
    cost = eng.ad_cost(gl[band.b0:band.b1].contiguous(), gr[band.b0:band.b1].contiguous(), D, 0)
    torch.cuda.synchronize()
    times, timings = [], {}
    out = None
    for r in range(a.reps):
        c = cost.clone()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = aggregate_cross_band(eng, bgr, c, band, iters=4, vmax=255.0, timings=timings)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        times.append(float(ms.item()))
    res = {"shape": [H, W, D], "world": world, "band_rows": band.r1 - band.r0, "ms_per_frame_max_over_ranks": min(times),
           "exchange_ms_rank0": timings.get("exchange_ms"), "exchanges": timings.get("exchanges"),
           "halo_MB_per_exchange_per_neighbour": timings.get("halo_bytes_per_exchange_per_neighbour", 0) / 1e6,
           "gdisp_evals_per_s": H * W * D / (min(times) / 1e3) / 1e9}
    checksum = torch.tensor([float(out.double().sum().item())], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(checksum)
    res["checksum"] = float(checksum.item())
    if a.verify != "none" and (a.verify == "all" or rank == 0):
        if rank == 0 or a.verify == "all":
            full = None
            if rank == 0:
                cfull = eng.ad_cost(gl, gr, D, 0)
                full = eng.aggregate_cross(bgr, cfull, mode=SUM_FAST)["cost"]
                del cfull
        if a.verify == "all" and world > 1:
            parts = [torch.empty((make_band(r, world, H).r1 - make_band(r, world, H).r0, W, D), device=dev) for r in range(world)] if rank == 0 else None
            dist.gather(out.contiguous(), parts, dst=0)
            if rank == 0:
                got = torch.cat(parts, 0)
        else:
            got, full = (out, full[band.r0:band.r1]) if rank == 0 else (None, None)
        if rank == 0:
            same = float((got == full).float().mean())
            rel = ((got - full).abs() / full.abs().clamp_min(1e-30))
            rel[got == full] = 0
            res.update(identical_fraction=same, max_rel_diff=float(rel.max()), full_frame_checksum=float(full.double().sum().item()) if a.verify == "all" else None)
    if rank == 0:
        print(json.dumps(res), flush=True)
    eng.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
