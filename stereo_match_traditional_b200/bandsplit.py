"""Row-band split of ONE very large frame over several GPUs (BASELINE.json configs[4]).

GPU g owns image rows [r0, r1) and keeps ``HALO`` = 34 (the maximum arm length, cross_aggregator.cpp:151 with the
reference's L1) extra rows on each interior side.  Everything that only needs the IMAGE is computed locally from the
full image (arms, support counts, and the cost of the band + halo rows: a cost row depends on image rows only, the
9x7 census needs +-4 of them), so the only data that must cross NVLink is the CURRENT COST of the 34 rows beyond each
band edge before every vertical aggregation pass: one neighbour exchange per iteration, 34*W*D*4 bytes per neighbour.
Horizontal passes are row-local.  The exchange is a pair of point-to-point transfers per neighbour
(``torch.distributed`` batch_isend_irecv: NCCL over NVLink on GPUs, gloo in the CPU tests); there is no collective.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

HALO = 34


@dataclass
class Band:
    rank: int
    world: int
    rows: int          # image rows
    r0: int            # first owned row
    r1: int            # one past the last owned row
    top: int           # halo rows above (0 for the first band)
    bottom: int        # halo rows below (0 for the last band)

    @property
    def b0(self) -> int:      # first buffered image row
        return self.r0 - self.top

    @property
    def b1(self) -> int:
        return self.r1 + self.bottom

    @property
    def nrows(self) -> int:   # buffered rows
        return self.b1 - self.b0


def make_band(rank: int, world: int, rows: int, halo: int = HALO) -> Band:
    """Even split of the rows; interior edges get `halo` extra rows."""
    per = (rows + world - 1) // world
    r0, r1 = min(rows, rank * per), min(rows, (rank + 1) * per)
    if r1 - r0 < halo and world > 1:
        raise ValueError(f"band of {r1 - r0} rows is shorter than the {halo}-row halo; use fewer GPUs")
    return Band(rank, world, rows, r0, r1, halo if rank > 0 else 0, halo if rank < world - 1 else 0)


def exchange_halos(buf, band: Band, group=None):
    """Refresh the halo rows of `buf` ([band.nrows, ...] tensor, rows contiguous) from the neighbours' owned rows.

    Sends my first / last `HALO` OWNED rows to the upper / lower neighbour and receives their edge rows into my halos."""
    import torch.distributed as dist

    ops = []
    t, b, n = band.top, band.bottom, band.nrows
    if band.rank > 0:  # upper neighbour: it needs my first owned rows as its bottom halo, I need its last owned rows
        ops.append(dist.P2POp(dist.isend, buf[t:t + HALO], band.rank - 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[0:t], band.rank - 1, group))
    if band.rank < band.world - 1:
        ops.append(dist.P2POp(dist.isend, buf[n - b - HALO:n - b], band.rank + 1, group))
        ops.append(dist.P2POp(dist.irecv, buf[n - b:n], band.rank + 1, group))
    if ops:
        for r in dist.batch_isend_irecv(ops):
            r.wait()
    return buf


def aggregate_cross_band(engine, bgr_full, cost_band, band: Band, iters: int = 4, L1=34, L2=17, t1=20, t2=6, vmax=None, group=None,
                         timings: dict | None = None, mode: int = 1):
    """CrossAggregator::Aggregate(iters) on this rank's band.

    bgr_full: uint8 [H, W, 3] CUDA tensor of the WHOLE image (arms need +-34 rows of it).
    cost_band: float32 [band.nrows, W, D] CUDA tensor, the initial cost of the buffered rows (owned + halo).
    mode: 1 = SMT_SUM_FAST (default), 0 = SMT_SUM_REFERENCE_ORDER (bands bit-identical to the whole frame).
    Returns the aggregated cost of the OWNED rows, [r1-r0, W, D] (a view into one of two band buffers)."""
    import torch

    if L1 > HALO:
        raise ValueError("the band halo assumes arms <= 34 (cross_L1 <= 34)")
    H, W = bgr_full.shape[:2]
    D = cost_band.shape[2]
    lib, h = engine.lib, engine.h
    maps = engine.cross_arms(bgr_full, L1, L2, t1, t2)        # full-image maps: integer work on 100 MB at most
    arms = maps["arms"][band.b0:band.b1].contiguous()
    sup0 = maps["sup0"][band.b0:band.b1].contiguous()
    sup1 = maps["sup1"][band.b0:band.b1].contiguous()
    if vmax is None:
        vmax = float(cost_band.max().item()) if float(cost_band.min().item()) >= 0 else -1.0
    cur, oth = cost_band, torch.empty_like(cost_band)
    nsweeps = lib.smt_cross_sweep_count(iters)
    ev = []
    for s in range(nsweeps):
        horiz, div = C.c_int32(0), C.c_int32(0)
        lib.smt_cross_sweep_info(iters, s, C.byref(horiz), C.byref(div))
        if not horiz.value and band.world > 1:
            engine.sync()                                   # the previous sweep wrote `cur`
            if timings is not None:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
            exchange_halos(cur, band, group)
            if timings is not None:
                e1.record()
                ev.append((e0, e1))
            torch.cuda.current_stream().synchronize()       # halos landed before the engine's stream reads them
        engine._ck(lib.smt_aggregate_cross_sweep(h, iters, s, C.c_void_p(cur.data_ptr()), C.c_void_p(oth.data_ptr()),
                                                 C.c_void_p(arms.data_ptr()), C.c_void_p(sup0.data_ptr()),
                                                 C.c_void_p(sup1.data_ptr()), band.nrows, W, D, float(vmax), int(mode), 1))
        cur, oth = oth, cur
    engine.sync()
    if timings is not None:
        timings["exchange_ms"] = sum(a.elapsed_time(b) for a, b in ev)
        timings["exchanges"] = len(ev)
        timings["halo_bytes_per_exchange_per_neighbour"] = HALO * W * D * 4
    return cur[band.top:band.top + (band.r1 - band.r0)]
