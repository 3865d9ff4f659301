"""Row-band split of ONE very large frame over several GPUs (BASELINE.json configs[4]).

GPU g owns image rows [r0, r1) and keeps ``HALO`` = 34 (the maximum arm length, cross_aggregator.cpp:151 with the
reference's L1) extra rows on each interior side.  Everything that only needs the IMAGE is computed locally from the
full image (arms, support counts, the cost of the owned rows), so the only data that must cross NVLink is the CURRENT
COST of the 34 rows beyond each band edge before every vertical aggregation pass.

The product path is native: ``libsmt_b200.so`` keeps a band in one device allocation that the neighbours map (CUDA IPC),
pushes edge rows straight into the neighbours' halo rows with copy engines, and orders the passes with sequence numbers
in device memory (csrc/bandsplit.cu; ``include/smt_b200.h``: smt_band_*, smt_halo_exchange).  This module is its host
side: the ``DeviceBand`` wrapper, the handle exchange over ``torch.distributed`` (a few hundred bytes, once), and
``aggregate_cross_band`` which puts them together.

``make_band`` / ``exchange_halos`` are the same partition and the same exchange pattern on ``torch.distributed``
point-to-point transfers; the CPU tests (gloo, tests/test_bandsplit_host.py) run them where no GPU is present.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

HALO = 34
HANDLE_BYTES = 128


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


def make_band(rank: int, world: int, rows: int) -> Band:
    """Even split of the rows; interior edges get HALO extra rows.  The split is a pure function of (rows, world): a
    band shorter than the halo ANYWHERE raises on EVERY rank, so that no rank enters an exchange its neighbour skipped."""
    per = (rows + world - 1) // world
    for r in range(world):
        a, e = min(rows, r * per), min(rows, (r + 1) * per)
        if world > 1 and e - a < HALO:
            raise ValueError(f"band {r} of {world} would own {e - a} rows, fewer than the {HALO}-row halo; use fewer GPUs")
    r0, r1 = min(rows, rank * per), min(rows, (rank + 1) * per)
    return Band(rank, world, rows, r0, r1, HALO if rank > 0 else 0, HALO if rank < world - 1 else 0)


def exchange_halos(buf, band: Band, group=None):
    """Host-transport form of the exchange (torch.distributed point-to-point; gloo in the CPU tests): refresh the halo rows
    of `buf` ([band.nrows, ...] tensor) from the neighbours' owned edge rows.  Extents come from the band, so both ends
    of every transfer agree on its size."""
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


class DeviceBand:
    """One rank's band in device memory (smt_band).  `volume` is a torch view of volume 0, [nrows, cols, D] float32: write
    the initial cost of the OWNED rows at rows [top, top + owned), read the aggregated cost from the same rows."""

    def __init__(self, engine, image_rows: int, cols: int, D: int, rank: int, world: int):
        import torch

        self.engine, self.lib = engine, engine._raw
        self.h = C.c_void_p()
        engine._ck(self.lib.smt_band_create(engine.h, image_rows, cols, D, rank, world, C.byref(self.h)))
        r0, r1, top, bottom = (C.c_int32() for _ in range(4))
        self.lib.smt_band_geometry(self.h, C.byref(r0), C.byref(r1), C.byref(top), C.byref(bottom))
        self.geom = Band(rank, world, image_rows, r0.value, r1.value, top.value, bottom.value)
        self.cols, self.D = cols, D
        ptr = C.c_void_p()
        engine._ck(self.lib.smt_band_volume(self.h, C.byref(ptr)))
        n = self.geom.nrows * cols * D

        class _Mem:  # __cuda_array_interface__ view of the library-owned allocation
            __cuda_array_interface__ = {"shape": (n,), "typestr": "<f4", "data": (ptr.value, False), "version": 3}

        self._mem = _Mem()
        with torch.cuda.device(engine.device):
            self.volume = torch.as_tensor(self._mem, device=f"cuda:{engine.device}").view(self.geom.nrows, cols, D)

    @property
    def owned(self):
        g = self.geom
        return self.volume[g.top:g.top + (g.r1 - g.r0)]

    def export(self) -> bytes:
        buf = (C.c_ubyte * HANDLE_BYTES)()
        self.engine._ck(self.lib.smt_band_export(self.h, buf))
        return bytes(buf)

    def connect(self, upper: bytes | None, lower: bytes | None):
        def arg(b):
            return None if b is None else (C.c_ubyte * HANDLE_BYTES).from_buffer_copy(b)

        self._keep = (arg(upper), arg(lower))
        self.engine._ck(self.lib.smt_band_connect(self.h, self._keep[0], self._keep[1]))

    def connect_local(self, upper: "DeviceBand | None", lower: "DeviceBand | None"):
        self.engine._ck(self.lib.smt_band_connect_local(self.h, upper.h if upper else None, lower.h if lower else None))

    def connect_group(self, group=None):
        """All-gather the handles over torch.distributed (once) and map the two neighbours."""
        import torch.distributed as dist

        g = self.geom
        if g.world == 1:
            return
        handles = [None] * g.world
        dist.all_gather_object(handles, self.export(), group=group)
        self.connect(handles[g.rank - 1] if g.rank > 0 else None, handles[g.rank + 1] if g.rank < g.world - 1 else None)
        dist.barrier(group=group)   # every rank has mapped its neighbours before anyone pushes

    def halo_exchange(self):
        self.engine._ck(self.lib.smt_halo_exchange(self.h))

    def aggregate_cross(self, arms, sup0, sup1, iters=4, vmax=-1.0, mode=1):
        """arms [nrows, cols, 4] u8, sup0 / sup1 [nrows, cols] u16-as-int16 CUDA tensors: the band's rows of the full maps."""
        self._maps = (arms, sup0, sup1)   # alive until the asynchronous passes have run
        self.engine.order_after_torch()
        self.engine._ck(self.engine._raw.smt_band_aggregate_cross(self.h, C.c_void_p(arms.data_ptr()), C.c_void_p(sup0.data_ptr()),
                                                                  C.c_void_p(sup1.data_ptr()), int(iters), float(vmax), int(mode)))
        # torch's stream is NOT queued behind the band: the band's stream waits for its neighbours, and a caller that drives
        # several bands from one torch stream would chain them into a deadlock.  Results are valid after sync().

    def sync(self) -> float:
        """Wait for the band's streams; raises if an exchange timed out.  Returns the bytes this rank has pushed."""
        v = C.c_double(0.0)
        self.engine._ck(self.lib.smt_band_sync(self.h, C.byref(v)))
        return float(v.value)

    def close(self):
        if self.h:
            self.lib.smt_band_destroy(self.h)
            self.h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def band_maps(engine, bgr_full, band: Band, L1=34, L2=17, t1=20, t2=6):
    """Arms and support counts of the band's buffered rows, computed from the FULL image (vertical arms need +-34 rows
    of it; integer work on an image-sized array)."""
    if L1 > HALO:
        raise ValueError("the band halo assumes arms <= 34 (cross_L1 <= 34)")
    maps = engine.cross_arms(bgr_full, L1, L2, t1, t2)
    return tuple(maps[k][band.b0:band.b1].contiguous() for k in ("arms", "sup0", "sup1"))


def aggregate_cross_band(engine, bgr_full, cost_owned, dband: DeviceBand, iters: int = 4, L1=34, L2=17, t1=20, t2=6, vmax=None, mode: int = 1):
    """CrossAggregator::Aggregate(iters) on this rank's band (connected DeviceBand).

    bgr_full: uint8 [H, W, 3] CUDA tensor of the WHOLE image.  cost_owned: float32 [r1-r0, W, D] CUDA tensor, the initial
    cost of the owned rows.  mode: 1 = SMT_SUM_FAST, 0 = SMT_SUM_REFERENCE_ORDER (bands bit-identical to the whole frame).
    Returns the aggregated cost of the owned rows (a view into the band; valid after dband.sync())."""
    arms, sup0, sup1 = band_maps(engine, bgr_full, dband.geom, L1, L2, t1, t2)
    if vmax is None:
        vmax = float(cost_owned.max().item()) if float(cost_owned.min().item()) >= 0 else -1.0
    dband.owned.copy_(cost_owned)
    dband.aggregate_cross(arms, sup0, sup1, iters, vmax, mode)
    return dband.owned
