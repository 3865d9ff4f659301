"""bench.py --workload: the BASELINE.json configurations other than the headline (configs[0], [1], [2], [4]).

Same contract as bench.py's default workload: W >= 3 warm-up steps, exactly K timed steps between barriers and
synchronisations, CUDA events, one JSON line with `value` (inputs resident on the device), `e2e` (host buffers through the
host-memspace C ABI, copies inside the timed region) and `cpu_baseline` (the compiled reference -- or the C port -- on the
box's host cores, on a bounded sample that is described in the line).  Inputs are larger than nothing in particular here:
these frames fit in L2, so the steps rotate through distinct pairs and `config.l2` says so.
"""
from __future__ import annotations

import json
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _oracle():
    from oracle import oracle as orc

    kind = "ref" if orc.have("ref") else "port"
    if kind == "port" and not orc.have("port"):
        orc.build("port")
    return orc.Oracle(kind), ("reference" if kind == "ref" else "port")


def _timed(torch, stream, fn, steps):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            fn()
        e1.record(stream)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)


def run_workload(args):
    import torch

    from bench import ClockSampler, host_cores
    from stereo_match_traditional_b200.api import SUM_FAST, SUM_REFERENCE_ORDER, Engine
    from stereo_match_traditional_b200.synth import make_pair, replicate_pad

    if args.workload == "bandsplit":
        from scripts.bench_bandsplit import run_bandsplit

        return run_bandsplit(args)
    if int(os.environ.get("RANK", "0")) != 0:
        return
    name = args.workload
    steps, warmup = args.steps, max(3, args.warmup)
    B = 8  # distinct pairs per step
    if name == "sad_teddy":
        H, W, D = 375, 450, 64
        metric = "SAD 9x9 + WTA frames/s at 450x375 D=64 (BASELINE.json configs[0])"
        taps = 81
    elif name == "pfull_teddy":
        H, W, D = 375, 450, 64
        metric = "AD-Census P_full frames/s at 450x375 D=64 (BASELINE.json configs[1])"
        taps = None
    else:
        H, W, D = 720, 1280, 128
        metric = "NCC 21x21 + ASW 25x25 (left, right, cross-check) frames/s at 1280x720 D=128 (BASELINE.json configs[2])"
        taps = 441 * 5 + 625 * 2
        B = 2
    pairs = [make_pair(H, W, D, 1001 + k, "scene", frame=k) for k in range(B)]

    if args.impl == "reference":
        o, kind = _oracle()
        p = pairs[0]
        t0 = time.perf_counter()
        if name == "sad_teddy":
            lp, rp = replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)
            o.sad(lp, rp, D, 3)
            frac, what = 1.0, "one whole frame"
        elif name == "pfull_teddy":
            L, R = p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32)
            c = o.adcensus(L, R, D, want=("costL", "costR"))
            aL = o.cross_aggregate(p["bgr_left"], p["bgr_right"], c["costL"])["cost"]
            aR = o.cross_aggregate(p["bgr_right"], p["bgr_left"], c["costR"])["cost"]
            sl = o.scanline(aL, L, 10, 150, want=("disp",))["disp"]
            lrc, _, _ = o.lrcheck(sl, o.wta(aR), 2)
            o.median(o.remove_speckles(lrc, 1, 30, None), 3)
            frac, what = 1.0, "one whole frame"
        else:
            nr = 4
            r0 = 300
            o.ncc(p["gray_left"][r0 - 10:r0 + nr + 10], p["gray_right"][r0 - 10:r0 + nr + 10], D, 10)
            lp, rp = replicate_pad(p["gray_left"], 12), replicate_pad(p["gray_right"], 12)
            sl = slice(r0, r0 + nr + 24)
            o.asw(lp[sl], rp[sl], D, 11, 50, 30, 40)
            o.asw(lp[sl], rp[sl], D, 11, 50, 30, 40, True)
            frac, what = nr / float(H), f"{nr} output rows of one frame (NCC, ASW left and right), scaled by rows/{H}"
        dt = time.perf_counter() - t0
        fps = frac / dt
        print(json.dumps({"impl": "reference", "metric": metric, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": 1, "warmup": 0,
                          "ms_per_step": 1e3 * dt, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                          "dtype": "f64" if name == "ncc_asw_720p" else ("i32" if name == "sad_teddy" else "f32"), "data": "synthetic",
                          "config": {"workload": name, "shape": f"{H}x{W}x{D}"},
                          "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": 1, "kind": kind, "sample": what + f"; {dt:.1f} s"},
                          "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}),
              flush=True)
        return

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    stream = torch.cuda.Stream(device=dev)
    eng = Engine(0, stream.cuda_stream)
    if name == "sad_teddy":
        host = [(replicate_pad(p["gray_left"], 4), replicate_pad(p["gray_right"], 4)) for p in pairs]
        devin = [tuple(torch.as_tensor(a, device=dev) for a in h) for h in host]
        run_dev = lambda k: eng.sad(devin[k][0], devin[k][1], D, 3)
        run_host = lambda k: eng.sad(host[k][0], host[k][1], D, 3)
        h2d, d2h = 2 * host[0][0].size, H * W * 4
    elif name == "pfull_teddy":
        keys = ("bgr_left", "bgr_right", "gray_left", "gray_right")
        devin = [{k: torch.as_tensor(p[k], device=dev) for k in keys} for p in pairs]
        params = eng.default_params(H, W, D)
        params.sum_mode = SUM_FAST if args.sum_mode == "fast" else SUM_REFERENCE_ORDER
        run_dev = lambda k: eng.pipeline(devin[k]["bgr_left"], devin[k]["bgr_right"], D, devin[k]["gray_left"], devin[k]["gray_right"], params=params)
        run_host = lambda k: eng.pipeline(pairs[k]["bgr_left"], pairs[k]["bgr_right"], D, pairs[k]["gray_left"], pairs[k]["gray_right"], params=params)
        h2d, d2h = sum(pairs[0][k].size for k in keys), H * W * 4
    else:
        host = [(p["gray_left"], p["gray_right"], replicate_pad(p["gray_left"], 12), replicate_pad(p["gray_right"], 12)) for p in pairs]
        devin = [tuple(torch.as_tensor(a, device=dev) for a in h) for h in host]

        def both(src, k):
            gl, gr, lp, rp = src[k]
            eng.ncc(gl, gr, D, 10)
            dl = eng.asw(lp, rp, D, 11, 50, 30, 40)
            dr = eng.asw(lp, rp, D, 11, 50, 30, 40, right_view=True)
            return eng.asw_crosscheck(dl, dr)

        run_dev = lambda k: both(devin, k)
        run_host = lambda k: both(host, k)
        h2d, d2h = host[0][0].size * 2 + host[0][2].size * 4, H * W * (4 + 4 + 4 + 1)

    def step_dev():
        with torch.cuda.stream(stream):
            for k in range(B):
                run_dev(k)

    def step_host():
        for k in range(B):
            run_host(k)

    for _ in range(warmup):
        step_dev()
    eng.sync()
    l0 = eng.kernel_launches
    with ClockSampler(0) as clk:
        ms = _timed(torch, stream, step_dev, steps)
    launches = eng.kernel_launches - l0
    step_host()
    ms_e2e = _timed(torch, stream, step_host, steps)
    fps = B * steps / (ms / 1e3)
    line = {"metric": metric, "value": fps, "unit": "frames/s", "n_gpus": 1, "steps": steps, "warmup": warmup, "ms_per_step": ms / steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if name == "ncc_asw_720p" else ("i32" if name == "sad_teddy" else "f32"), "data": "synthetic",
            "gdisp_evals_per_s": fps * H * W * D / 1e9,
            "config": {"workload": name, "shape": f"{H}x{W}x{D}", "pairs_per_step": B,
                       "l2": f"{B} distinct pairs per step; a {H}x{W} pair and its outputs fit in L2, so this line measures the L2-resident regime"},
            "clocks": clk.summary(),
            "e2e": {"value": B * steps / (ms_e2e / 1e3), "unit": "frames/s", "h2d_bytes_per_step": int(h2d) * B, "d2h_bytes_per_step": int(d2h) * B},
            "gpu_launches": int(launches)}
    if taps:
        line["tap_evals_per_s"] = fps * H * W * D * taps
    if not args.no_cpu_baseline:
        import subprocess

        r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--workload", name, "--impl", "reference"], capture_output=True, text=True,
                           cwd=str(ROOT))
        try:
            line["cpu_baseline"] = json.loads(r.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception:
            line["cpu_baseline"] = {"value": None, "unit": "frames/s", "cores": 1, "kind": "unavailable", "sample": r.stderr[-200:]}
    print(json.dumps(line), flush=True)
    eng.close()
