"""CPU, world_size 2 and 3 over gloo: band partition and the halo exchange pattern of the row-band split."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from stereo_match_traditional_b200.bandsplit import HALO, exchange_halos, make_band


def test_band_partition():
    rows = 4320
    bands = [make_band(r, 8, rows) for r in range(8)]
    assert bands[0].r0 == 0 and bands[-1].r1 == rows
    assert all(a.r1 == b.r0 for a, b in zip(bands, bands[1:]))
    assert bands[0].top == 0 and bands[0].bottom == HALO and bands[3].top == HALO and bands[-1].bottom == 0
    assert bands[3].nrows == 540 + 2 * HALO
    with pytest.raises(ValueError):
        make_band(1, 8, 100)          # 13-row bands cannot hold a 34-row halo
    # rows=290, world=8: ranks 0-6 would own 37 rows and only rank 7 31 -- every rank must refuse, or 0-6 would wait for 7
    for r in range(8):
        with pytest.raises(ValueError):
            make_band(r, 8, 290)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, rows, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    band = make_band(rank, world, rows)
    # every buffered row holds its global row index in the owned part and -1 in the halos
    buf = torch.full((band.nrows, 3, 2), -1.0)
    buf[band.top:band.top + band.r1 - band.r0] = torch.arange(band.r0, band.r1, dtype=torch.float32)[:, None, None]
    exchange_halos(buf, band)
    want = torch.arange(band.b0, band.b1, dtype=torch.float32)[:, None, None].expand_as(buf)
    q.put((rank, bool(torch.equal(buf, want))))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_fills_halos_with_neighbour_rows(world):
    port, rows = _free_port(), 150
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, rows, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(res[r] for r in range(world)), res
