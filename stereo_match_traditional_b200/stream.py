"""Frame-stream sharding for multi-GPU runs (BASELINE.json configs[3]: batch-parallel, no collective).

Stereo pairs are independent, so a stream of frames is dealt to ranks round-robin in blocks of `frames_per_step`; the
only cross-rank operations are the barrier around the timed region and a MAX reduction of the per-rank device times
(NCCL on GPUs, gloo in the CPU tests)."""
from __future__ import annotations


def frames_for_rank(rank: int, world: int, frames_per_step: int, step: int = 0) -> list[int]:
    """Global frame ids rank `rank` processes in step `step`: ranks own disjoint, contiguous blocks of each step."""
    base = step * world * frames_per_step + rank * frames_per_step
    return list(range(base, base + frames_per_step))


def frame_seed(frame_id: int) -> int:
    """seed 1004 + frame for the 1920x1080 stream."""
    return 1004 + frame_id


def reduce_max(value_ms: float, device=None) -> float:
    """Max over ranks of a per-rank time (identity without an initialised process group)."""
    import torch
    import torch.distributed as dist

    if not (dist.is_available() and dist.is_initialized()):
        return float(value_ms)
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def aggregate_fps(world: int, frames_per_step: int, steps: int, max_ms: float) -> float:
    """Whole-job frames/s: all frames of all ranks over the slowest rank's time."""
    return world * frames_per_step * steps / (max_ms / 1e3)
