#!/usr/bin/env python
"""Benchmark of the AD-Census full pipeline (P_full) -- BASELINE.json's metric:
"AD-Census frames/s and Gdisp-evals/s (W*H*D/s) at 1920x1080 D=256".

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

One step = one pass of P_full over a batch of `--frames-per-step` synthetic rectified pairs per GPU, taken in turn from a
stream of 64 distinct pairs per GPU (config 4 of BASELINE.json: batched frame stream, batch-parallel over GPUs, no
collective).  Prints ONE JSON line.  --workload selects the other BASELINE.json configurations (sad_teddy, pfull_teddy,
ncc_asw_720p, bandsplit); the default is the headline, pfull_1080p.

  value     frames/s over all ranks, inputs already resident in HBM, CUDA events on the launching stream,
            barrier + synchronize on both sides, max over ranks
  e2e       the same metric through the host-memspace C-ABI call (pinned host images in, disparity map out; the
            H2D / D2H copies are inside the timed region)
  roofline  dominant stage's kernel: algorithmic bytes / CUDA-event time vs the measured HBM peak
  cpu_baseline / --impl reference
            the reference's own CPU code (oracle/_ref: the unmodified sources compiled with g++ -O3; falls back to
            the C port) on a bounded sample of the same workload -- full-width row bands of the same frames, one
            independent stream per host core (the reference's AD-Census path has no threading of its own)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

import numpy as np  # noqa: E402

METRIC = "AD-Census frames/s at 1920x1080 D=256 (Gdisp-evals/s = W*H*D*fps/1e9 in `gdisp_evals_per_s`)"
STREAM_FRAMES = 64   # distinct frames per GPU the steps rotate through (cfg 4)


# ----------------------------------------------------------------------------------------------- CPU reference arm
def cpu_worker(rows: int, cols: int, D: int, seed: int, reps: int, full_rows: int = 0) -> None:
    """Child process: P_full with the reference's own code on a band of `rows` rows of a `full_rows` x cols frame (or on
    a whole rows x cols frame when full_rows is 0); prints seconds per repetition."""
    from oracle import oracle as orc
    from stereo_match_traditional_b200.synth import make_pair

    kind = "ref" if orc.have("ref") else "port"
    if kind == "port" and not orc.have("port"):
        orc.build("port")
    o = orc.Oracle(kind)
    if full_rows:
        full = make_pair(full_rows, cols, D, seed, "scene")
        r0 = (seed * 97) % max(1, full_rows - rows)
        p = {k: np.ascontiguousarray(v[r0:r0 + rows]) for k, v in full.items()}
    else:
        p = make_pair(rows, cols, D, seed, "scene")
    L = p["gray_left"].astype(np.float32)
    R = p["gray_right"].astype(np.float32)
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        c = o.adcensus(L, R, D, want=("costL", "costR"))
        aggL = o.cross_aggregate(p["bgr_left"], p["bgr_right"], c["costL"])["cost"]
        aggR = o.cross_aggregate(p["bgr_right"], p["bgr_left"], c["costR"])["cost"]
        sl = o.scanline(aggL, L, 10, 150, want=("disp",))["disp"]
        lrc, _, _ = o.lrcheck(sl, o.wta(aggR), 2)
        o.median(o.remove_speckles(lrc, 1, 30, None), 3)
        times.append(time.perf_counter() - t0)
    print(json.dumps({"kind": "reference" if kind == "ref" else "port", "times": times}))


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def run_cpu_streams(rows: int, cols: int, D: int, reps: int, workers: int, full_rows: int = 0):
    """`workers` independent frame streams (one process each).  Returns (kind, wall seconds per step list)."""
    procs = []
    for w in range(workers):
        cmd = [sys.executable, str(ROOT / "bench.py"), "--_cpu_worker", f"{rows},{cols},{D},{1004 + w},{reps},{full_rows}"]
        procs.append(subprocess.Popen(cmd, stdout=subprocess.PIPE, text=True, cwd=str(ROOT)))
    outs = []
    for p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError("cpu worker failed")
        outs.append(json.loads(out.strip().splitlines()[-1]))
    kind = outs[0]["kind"]
    # step r finishes when the slowest stream finishes repetition r
    per_step = [max(o["times"][r] for o in outs) for r in range(reps)]
    return kind, per_step


def cpu_sample_rows(H, W, D, budget_s: float) -> int:
    """Rows of a FULL-WIDTH band of the workload's frame whose reference run takes about budget_s seconds on one core
    (~1.8 Mcell/s measured on the GPU box).  Full width keeps the horizontal arms as long as in the whole frame; the
    vertical arms of a band shorter than 69 rows are cut at its borders, which makes the sample cheaper per cell than the
    whole frame, i.e. the scaled figure errs in the CPU's favour."""
    cells = 1.8e6 * budget_s
    return int(max(8, min(H, cells / (W * D))))


def workload_config(H, W, D, B, lanes, sum_mode, world):
    """`config` of the JSON line: the same object in both arms (the CPU arm's sample is described in cpu_baseline.sample)."""
    return {"workload": f"AD-Census P_full {W}x{H} D={D} (BASELINE.json configs[3]): census 9x7 + AD cost, cross aggregation x4 "
                        f"both views, 4-path scanline, WTA, LR check, speckles, median; stream of {STREAM_FRAMES} distinct frames per GPU",
            "frames_per_step_per_gpu": B, "streams_per_gpu": lanes, "sum_mode": sum_mode,
            "parallelism": f"frame-parallel x{world}, no collective",
            "l2": "no explicit flush: every frame streams two 2.1 GB cost volumes (>> 126 MB L2) and consecutive frames differ"}


def reference_arm(args, H, W, D):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    workers = min(host_cores(), 32)
    budget = max(1.0, 150.0 / (args.steps + args.warmup))
    rows = cpu_sample_rows(H, W, D, budget)
    kind, per_step = run_cpu_streams(rows, W, D, args.steps + args.warmup, workers, full_rows=H)
    timed = per_step[args.warmup:]
    total = sum(timed)
    frac = rows / float(H)  # fraction of a full frame each sample is
    fps = workers * frac * len(timed) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(timed), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "gdisp_evals_per_s": fps * H * W * D / 1e9,
        "config": workload_config(H, W, D, args.frames_per_step, max(1, args.streams), args.sum_mode, max(1, args.gpus)),
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": workers, "kind": kind,
                         "sample": f"per step and core one full-width {W}x{rows} band (rows taken from the {W}x{H} frames of the "
                                   f"stream, D={D}), full P_full on it (both views), scaled by rows/{H}; {workers} independent "
                                   "streams, because the reference's AD-Census path is single-threaded; vertical arms are cut at "
                                   "the band borders, which errs in the CPU's favour"},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.stop = threading.Event()
        self.thread = None

    def _run_nvml(self):
        """NVML polling (25 ms): enough samples inside a sub-second timed region.  Rows use the nvidia-smi layout."""
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(self.index)
        mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        bits = [("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4)]
        while not self.stop.is_set():
            sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            r = int(get_reasons(h))
            self.rows.append([str(sm), str(mx), "", hex(r)] + [("Active" if r & b else "Not Active") for _, b in bits])
            self.stop.wait(0.025)
        nv.nvmlShutdown()

    def _run(self):
        try:
            self._run_nvml()
            return
        except Exception:
            pass
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.2)

    def __enter__(self):
        self.thread = threading.Thread(target=self._run, daemon=True)
        self.thread.start()
        return self

    def __exit__(self, *exc):
        self.stop.set()
        self.thread.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for n, v in zip(names, r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]), "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------------- GPU arm
def _make_frame(a):
    from stereo_match_traditional_b200.stream import frame_seed
    from stereo_match_traditional_b200.synth import make_pair

    H, W, D, f = a
    return make_pair(H, W, D, frame_seed(f), "scene", frame=f)


def make_frames(H, W, D, ids):
    """The synthetic pairs of the given global frame ids (host arrays), generated on the host's cores."""
    import multiprocessing as mp

    n = min(len(ids), host_cores(), 16)
    if n <= 1:
        return [_make_frame((H, W, D, f)) for f in ids]
    with mp.get_context("fork").Pool(n) as pool:
        return pool.map(_make_frame, [(H, W, D, f) for f in ids])


def measured_hbm_peak():
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, STREAM-style copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def gpu_arm(args, H, W, D):
    import torch
    import torch.distributed as dist

    from stereo_match_traditional_b200.api import SUM_FAST, SUM_REFERENCE_ORDER, Engine
    from stereo_match_traditional_b200.stream import aggregate_fps, reduce_max
    from stereo_match_traditional_b200.synth import make_pair

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.frames_per_step
    # a stream of distinct frames per rank (seed 1004 + frame); the steps walk through it B at a time.
    # Generated first: the worker pool forks, which must happen before this process touches CUDA.
    nstream = max(B, (args.stream_frames // B) * B)
    frames = make_frames(H, W, D, [rank * nstream + f for f in range(nstream)])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    mode = SUM_FAST if args.sum_mode == "fast" else SUM_REFERENCE_ORDER

    keys = ("bgr_left", "bgr_right", "gray_left", "gray_right")
    dev_frames = [{k: torch.as_tensor(fr[k], device=dev) for k in keys} for fr in frames]
    pin_frames = [{k: torch.as_tensor(fr[k]).pin_memory() for k in keys} for fr in frames]
    pin_out = [torch.empty((H, W), dtype=torch.float32).pin_memory() for _ in range(B)]
    h2d = sum(int(t.numel()) * t.element_size() for t in pin_frames[0].values()) * B
    d2h = H * W * 4 * B
    cursor = {"dev": 0, "host": 0}

    def next_batch(which):
        k = cursor[which]
        cursor[which] = (k + B) % nstream
        return range(k, k + B)

    stream = torch.cuda.Stream(device=dev)
    eng = Engine(local, stream.cuda_stream)
    params = eng.default_params(H, W, D)
    params.sum_mode = mode
    # --streams 2: the frames of a step alternate between two contexts on two streams (independent stereo pairs, the way a
    # throughput-oriented caller would drive the library); the timed region starts and ends on `stream`
    extra = []
    for _ in range(max(1, args.streams) - 1):
        st2 = torch.cuda.Stream(device=dev)
        extra.append((st2, Engine(local, st2.cuda_stream)))
    lanes = [(stream, eng)] + extra

    def step_device():
        for st2, _ in extra:
            st2.wait_stream(stream)
        for k, f in enumerate(next_batch("dev")):
            fr = dev_frames[f]
            st_k, eng_k = lanes[k % len(lanes)]
            with torch.cuda.stream(st_k):
                eng_k.pipeline(fr["bgr_left"], fr["bgr_right"], D, fr["gray_left"], fr["gray_right"], params=params)
        for st2, _ in extra:
            stream.wait_stream(st2)

    host_pool = None
    if len(lanes) > 1:
        from concurrent.futures import ThreadPoolExecutor

        host_pool = ThreadPoolExecutor(max_workers=len(lanes))   # one host thread per context: the C call releases the GIL

    def host_frame(eng_k, f, po):
        fr = pin_frames[f]
        return eng_k.pipeline(fr["bgr_left"].numpy(), fr["bgr_right"].numpy(), D, fr["gray_left"].numpy(), fr["gray_right"].numpy(),
                              params=params, out=po.numpy())["disp"]

    def step_host():
        batch = list(zip(next_batch("host"), pin_out))
        if host_pool is None:
            return [host_frame(eng, f, po) for f, po in batch]
        # the frames of a step alternate between the contexts, each driven by its own host thread, so that the copies of
        # one frame overlap the kernels of another; every call still copies in, runs, copies out and synchronises
        futs = [host_pool.submit(host_frame, lanes[k % len(lanes)][1], f, po) for k, (f, po) in enumerate(batch)]
        return [fu.result() for fu in futs]

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            for _ in range(steps):
                fn()
            e1.record(stream)
        barrier()
        return reduce_max(e0.elapsed_time(e1), dev)

    for _ in range(args.warmup):
        step_device()
    eng.sync()
    l0 = eng.kernel_launches
    l0x = [e2.kernel_launches for _, e2 in extra]
    with ClockSampler(local) as clk:
        ms = timed(step_device, args.steps)
    launches = eng.kernel_launches - l0 + sum(e2.kernel_launches - a for (_, e2), a in zip(extra, l0x))
    fps = aggregate_fps(world, B, args.steps, ms)

    # end-to-end through the host-memspace ABI (pinned host buffers, copies inside the timed region)
    step_host()
    ms_e2e = timed(step_host, args.steps)
    fps_e2e = aggregate_fps(world, B, args.steps, ms_e2e)

    # per-stage CUDA-event times for the roofline of the dominant kernel: frames one after the other on ONE context, so
    # that a stage's events bracket its own kernels only (with --streams 2 the other stream's kernels would run in between)
    eng.stage_timing(True)
    nst = max(1, min(args.steps, 3))
    with torch.cuda.stream(stream):
        for _ in range(nst):
            for f in next_batch("dev"):
                fr = dev_frames[f]
                eng.pipeline(fr["bgr_left"], fr["bgr_right"], D, fr["gray_left"], fr["gray_right"], params=params)
    st = {k: v / (nst * B) for k, v in eng.read_stage_timing().items()}  # ms per frame
    eng.stage_timing(False)
    V = 4.0 * H * W * D
    iters = params.agg_iters
    alg = {"cost": 2 * V, "aggregate": 2 * V * iters * 2, "scanline": 2 * V, "wta": 1 * V}   # SURVEY.md 8(d), both views
    # the fast aggregation runs its 8 passes per view as 5 sweeps (H | V'V | H'H | V'V | H'): 3 fused-pair launches
    # (agg_fx2_kernel, two passes each) and 2 single-pass launches (agg_fx1_kernel); the roofline line is about the pair kernel,
    # "launch" = one sweep, algorithmic bytes per sweep = SURVEY 8(d)'s 1 V per pass x the passes the sweep covers
    kern = {"cost": "adcensus_cost_kernel", "aggregate": "agg_fx2_kernel (+ agg_fx1_kernel)" if mode == SUM_FAST else "agg_pass_exact44_kernel",
            "scanline": "scan_path_kernel", "wta": "wta_first_min_kernel"}
    nsweeps = (iters + 1) if iters > 0 else 0
    nlaunch = {"cost": 2, "aggregate": 2 * nsweeps if mode == SUM_FAST else 4 * iters, "scanline": 4, "wta": 1}  # per frame
    dom = max(alg, key=lambda k: st.get(k, 0.0))
    peak, how = measured_hbm_peak()
    ach = alg[dom] / (st[dom] / 1e3) / 1e9
    # dram__bytes_read + dram__bytes_write per launch of the dominant kernel from this round's `ncu --set full` capture
    # (profiles/r2_ncu_agg_fx.csv): a STORED figure, not re-measured by this run
    traffic, traffic_src = None, None
    tf = ROOT / "profiles" / "roofline_traffic.json"
    if tf.exists():
        try:
            tj = json.loads(tf.read_text())
            traffic = tj.get(kern[dom].split(" ")[0])
            traffic_src = tj.get("_source")
        except Exception:
            traffic = None
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src, "kernel": kern[dom], "peak_source": how,
                "algorithmic_bytes_per_launch": alg[dom] / nlaunch[dom], "avg_launch_ms": st[dom] / nlaunch[dom],
                "launches_per_frame": nlaunch[dom], "algorithmic_bytes_per_frame": alg[dom], "stage_ms_per_frame": st[dom],
                "all_stages": {k: {"ms_per_frame": st.get(k, 0.0), "algorithmic_GBps": (alg[k] / (st[k] / 1e3) / 1e9) if st.get(k) else None}
                               for k in alg},
                "other_stage_ms_per_frame": {k: v for k, v in st.items() if k not in alg},
                # SURVEY.md 8(d) asks for both denominators and for the whole frame's compulsory traffic (22 V)
                "peak_nominal": 8000.0, "frac_of_nominal": ach / 8000.0,
                "frame": {"algorithmic_bytes": sum(alg.values()), "achieved": sum(alg.values()) * fps / world / 1e9,
                          "frac": sum(alg.values()) * fps / world / 1e9 / peak}}

    line = {
        "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "gdisp_evals_per_s": fps * H * W * D / 1e9,
        "config": workload_config(H, W, D, B, len(lanes), args.sum_mode, world),
        "clocks": clk.summary(),
        "e2e": {"value": fps_e2e, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches),
        "roofline": roofline,
    }
    if world == 1 and mode == SUM_FAST:
        # the bit-exact aggregation mode (SMT_SUM_REFERENCE_ORDER) beside the benchmarked one, one timed step
        pex = eng.default_params(H, W, D)
        pex.sum_mode = SUM_REFERENCE_ORDER

        def step_exact():
            for fr in dev_frames[:B]:
                eng.pipeline(fr["bgr_left"], fr["bgr_right"], D, fr["gray_left"], fr["gray_right"], params=pex)

        step_exact()
        ms_ex = timed(step_exact, 1)
        line["reference_order_mode"] = {"value": aggregate_fps(world, B, 1, ms_ex), "unit": "frames/s", "ms_per_frame": ms_ex / B,
                                        "note": "same pipeline with the bit-exact aggregation (taps added in the reference's order)"}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        workers = min(host_cores(), 32)
        rows = cpu_sample_rows(H, W, D, 20.0)
        kind, per_step = run_cpu_streams(rows, W, D, 1, workers, full_rows=H)
        cpu_fps = workers * rows / float(H) / per_step[0]
        line["cpu_baseline"] = {"value": cpu_fps, "unit": "frames/s", "cores": workers, "kind": kind,
                                "sample": f"one full-width {W}x{rows} band of a {W}x{H} frame per core, D={D}, full P_full (both "
                                          f"views), scaled by rows/{H}; {per_step[0]:.1f} s wall"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    eng.close()
    for _, e2 in extra:
        e2.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--frames-per-step", type=int, default=2)
    ap.add_argument("--streams", type=int, default=2, help="contexts/streams (and host threads in the e2e leg) the frames of a step alternate between")
    ap.add_argument("--shape", default="1080x1920x256", help="HxWxD")
    ap.add_argument("--sum-mode", default="fast", choices=["fast", "reference_order"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--stream-frames", type=int, default=STREAM_FRAMES, help="distinct frames per GPU the steps rotate through")
    ap.add_argument("--workload", default="pfull_1080p", choices=["pfull_1080p", "sad_teddy", "pfull_teddy", "ncc_asw_720p", "bandsplit"])
    ap.add_argument("--_cpu_worker", default=None, help=argparse.SUPPRESS)
    args = ap.parse_args()
    if args._cpu_worker:
        r, c, d, s, n, fr = (int(x) for x in args._cpu_worker.split(","))
        cpu_worker(r, c, d, s, n, fr)
        return
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    H, W, D = (int(x) for x in args.shape.lower().split("x"))
    if args.workload != "pfull_1080p":
        from scripts.bench_workloads import run_workload   # the other BASELINE.json configurations
        run_workload(args)
    elif args.impl == "reference":
        reference_arm(args, H, W, D)
    else:
        gpu_arm(args, H, W, D)


if __name__ == "__main__":
    main()
