"""GPU: the native row-band split (BASELINE.json configs[4]; csrc/bandsplit.cu, smt_band_* / smt_halo_exchange).

Bands of one frame are aggregated with one-sided halo pushes between them and compared with the whole-frame aggregation
and with the oracle: several bands in ONE process on one GPU (smt_band_connect_local, every band its own context and
stream) and two PROCESSES sharing the GPU through CUDA IPC handles (smt_band_export / smt_band_connect) -- the same code
path the multi-GPU run takes (bench.py --workload bandsplit under torchrun), where only the peer mapping is remote."""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.bandsplit import HALO, DeviceBand, aggregate_cross_band, band_maps, make_band
from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _rel(got, ref):
    r = np.abs(got.astype(np.float64) - ref) / np.maximum(np.abs(ref.astype(np.float64)), 1e-30)
    r[got == ref] = 0
    return float(r.max())


def test_single_band_equals_whole_frame(engine):
    import torch

    H, W, D = 90, 160, 40
    p = make_pair(H, W, D, 51, "scene")
    bgr = torch.as_tensor(p["bgr_left"], device="cuda:0")
    gl, gr = torch.as_tensor(p["gray_left"], device="cuda:0"), torch.as_tensor(p["gray_right"], device="cuda:0")
    cost = engine.ad_cost(gl, gr, D, 0)
    for mode in (SUM_FAST, SUM_REFERENCE_ORDER):
        full = engine.aggregate_cross(bgr, cost, mode=mode)["cost"]
        with DeviceBand(engine, H, W, D, 0, 1) as band:
            out = aggregate_cross_band(engine, bgr, cost, band, iters=4, vmax=255.0, mode=mode)
            band.sync()
            if mode == SUM_REFERENCE_ORDER:
                assert torch.equal(out, full)
            else:
                assert _rel(out.cpu().numpy(), full.cpu().numpy()) <= 2e-6


@pytest.mark.parametrize("mode", [SUM_FAST, SUM_REFERENCE_ORDER], ids=["fast", "reference_order"])
@pytest.mark.parametrize("H,W,D,world,cost_kind", [(150, 120, 40, 2, "ad"), (330, 96, 33, 3, "ad"), (420, 70, 64, 4, "adcensus")])
def test_bands_in_one_process_match_whole_frame(engine, oracle, H, W, D, world, cost_kind, mode):
    import torch
    from stereo_match_traditional_b200.api import Engine

    p = make_pair(H, W, D, 52 + world, "scene")
    dev = "cuda:0"
    bgr = torch.as_tensor(p["bgr_left"], device=dev)
    gl, gr = torch.as_tensor(p["gray_left"], device=dev), torch.as_tensor(p["gray_right"], device=dev)
    if cost_kind == "ad":        # CBLSM ComputeAD: integers 0..255 (mostly redone on float pairs in the fast mode)
        cost, vmax = engine.ad_cost(gl, gr, D, 0), 255.0
    else:                        # AD-Census cost: stays on the fixed-point route
        cost, vmax = engine.adcensus_cost(gl.float(), gr.float(), D)["cost"], 2.0
    full = engine.aggregate_cross(bgr, cost, mode=mode)["cost"]
    engines = [Engine(0) for _ in range(world)]      # one context (stream, workspace) per band, like one per GPU
    bands = [DeviceBand(e, H, W, D, r, world) for r, e in enumerate(engines)]
    try:
        for r, b in enumerate(bands):
            b.connect_local(bands[r - 1] if r > 0 else None, bands[r + 1] if r < world - 1 else None)
        torch.cuda.synchronize()
        outs = []
        for rep in range(2):      # twice: sequence numbers and volume roles carry over between aggregations
            outs = [aggregate_cross_band(e, bgr, cost[b.geom.r0:b.geom.r1], b, iters=4, vmax=vmax, mode=mode) for e, b in zip(engines, bands)]
            pushed = [b.sync() for b in bands]
        got = torch.cat(outs, 0)
        row_bytes = W * D * 4
        assert pushed[0] == 2 * 4 * HALO * row_bytes          # first band: one neighbour, four exchanges, two aggregations
        if world > 2:
            assert pushed[1] == 2 * 4 * 2 * HALO * row_bytes
        if mode == SUM_REFERENCE_ORDER:
            assert torch.equal(got, full)        # band split == single GPU, bit for bit (test pyramid level 4)
        else:
            assert _rel(got.cpu().numpy(), full.cpu().numpy()) <= 2e-6
        ref = oracle.cross_aggregate(p["bgr_left"], p["bgr_right"], cost.cpu().numpy())["cost"]
        assert _rel(got.cpu().numpy(), ref) <= (0.0 if mode == SUM_REFERENCE_ORDER else 1e-5)
    finally:
        for b in bands:
            b.close()
        for e in engines:
            e.close()


def test_halo_exchange_alone(engine):
    """smt_halo_exchange: after it, every band's halo rows hold the neighbours' edge rows."""
    import torch
    from stereo_match_traditional_b200.api import Engine

    H, W, D, world = 200, 24, 8, 3
    engines = [Engine(0) for _ in range(world)]
    bands = [DeviceBand(e, H, W, D, r, world) for r, e in enumerate(engines)]
    try:
        for r, b in enumerate(bands):
            b.connect_local(bands[r - 1] if r > 0 else None, bands[r + 1] if r < world - 1 else None)
        for b in bands:
            g = b.geom
            b.volume.fill_(-1.0)
            b.owned.copy_(torch.arange(g.r0, g.r1, dtype=torch.float32, device="cuda:0")[:, None, None].expand(-1, W, D))
        torch.cuda.synchronize()
        for b in bands:
            b.halo_exchange()
        for b in bands:
            b.sync()
        for b in bands:
            g = b.geom
            want = torch.arange(g.b0, g.b1, dtype=torch.float32, device="cuda:0")[:, None, None].expand(-1, W, D)
            assert torch.equal(b.volume, want)
    finally:
        for b in bands:
            b.close()
        for e in engines:
            e.close()


def test_band_refuses_short_bands_on_every_rank(engine):
    from stereo_match_traditional_b200.api import SmtError

    for rank in range(8):        # rows=290, world=8: ranks 0-6 would own 37 rows, rank 7 only 31 -- all must fail alike
        with pytest.raises(SmtError):
            DeviceBand(engine, 290, 64, 8, rank, 8)
        with pytest.raises(ValueError):
            make_band(rank, 8, 290)


CHILD = r"""
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch, torch.distributed as dist
from stereo_match_traditional_b200.api import Engine, SUM_FAST, SUM_REFERENCE_ORDER
from stereo_match_traditional_b200.bandsplit import DeviceBand, aggregate_cross_band
from stereo_match_traditional_b200.synth import make_pair
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")                       # host channel for the 128-byte handles only
torch.cuda.set_device(0)
H, W, D = 240, 96, 32
p = make_pair(H, W, D, 77, "scene")
bgr = torch.as_tensor(p["bgr_left"], device="cuda:0")
gl, gr = torch.as_tensor(p["gray_left"], device="cuda:0"), torch.as_tensor(p["gray_right"], device="cuda:0")
ok = True
with Engine(0) as e, DeviceBand(e, H, W, D, rank, world) as band:
    band.connect_group()
    cost = e.adcensus_cost(gl.float(), gr.float(), D)["cost"]
    for mode in (SUM_REFERENCE_ORDER, SUM_FAST):
        full = e.aggregate_cross(bgr, cost, mode=mode)["cost"]
        g = band.geom
        out = aggregate_cross_band(e, bgr, cost[g.r0:g.r1], band, iters=4, vmax=2.0, mode=mode)
        band.sync()
        a, b = out.cpu().numpy(), full[g.r0:g.r1].cpu().numpy()
        if mode == SUM_REFERENCE_ORDER:
            ok = ok and bool(np.array_equal(a, b))
        else:
            r = np.abs(a.astype(np.float64) - b) / np.maximum(np.abs(b), 1e-30); r[a == b] = 0
            ok = ok and float(r.max()) <= 2e-6
        dist.barrier()
print("BAND_OK" if ok else "BAND_MISMATCH", rank, flush=True)
dist.destroy_process_group()
"""


def test_two_processes_share_halos_through_cuda_ipc():
    """Two ranks, one process each (both on cuda:0 here; one GPU each under torchrun): handles travel over gloo, the halo
    rows over peer-mapped device memory."""
    import subprocess

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    procs = []
    for rank in range(2):
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE="2")
        procs.append(subprocess.Popen([sys.executable, "-c", CHILD % str(ROOT)], env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                                      text=True))
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for rank, (p, o) in enumerate(zip(procs, outs)):
        assert p.returncode == 0 and f"BAND_OK {rank}" in o, o[-3000:]
