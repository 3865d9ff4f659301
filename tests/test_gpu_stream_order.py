"""GPU: calls with torch CUDA tensors are ordered against torch's current stream on both sides (the context runs on its
own non-blocking stream): inputs still being produced when the call is made, outputs consumed right after it."""
import numpy as np
import pytest

from stereo_match_traditional_b200.api import SUM_FAST
from stereo_match_traditional_b200.synth import make_pair

pytestmark = pytest.mark.gpu


def test_inputs_produced_late_on_torchs_stream(engine):
    import torch

    H, W, D = 96, 200, 32
    p = make_pair(H, W, D, 5, "scene")
    want = engine.adcensus_cost(p["gray_left"].astype(np.float32), p["gray_right"].astype(np.float32), D)["cost"]
    gl8 = torch.as_tensor(p["gray_left"], device="cuda:0")
    gr8 = torch.as_tensor(p["gray_right"], device="cuda:0")
    gl = torch.zeros((H, W), dtype=torch.float32, device="cuda:0")
    gr = torch.zeros((H, W), dtype=torch.float32, device="cuda:0")
    torch.cuda.synchronize()
    torch.cuda._sleep(200_000_000)              # ~0.1 s of queued work in front of the producers
    gl.copy_(gl8)
    gr.copy_(gr8)
    got = engine.adcensus_cost(gl, gr, D)["cost"]
    assert np.array_equal(got.cpu().numpy(), want)


def test_outputs_read_right_after_the_call_full_hd(engine):
    """1080p pipeline twice in a row with device tensors and no explicit sync: the second call's outputs reuse the blocks
    the first call's outputs freed; both must read back identically."""
    import torch

    H, W, D = 1080, 1920, 64
    p = make_pair(H, W, D, 1004, "scene")
    t = {k: torch.as_tensor(p[k], device="cuda:0") for k in ("bgr_left", "bgr_right", "gray_left", "gray_right")}
    names = ("disp", "left_raw", "right_raw")
    first = {n: v.cpu().numpy() for n, v in
             engine.pipeline(t["bgr_left"], t["bgr_right"], D, t["gray_left"], t["gray_right"], want=names, sum_mode=SUM_FAST).items()}
    engine.sync()
    for _ in range(3):
        r = engine.pipeline(t["bgr_left"], t["bgr_right"], D, t["gray_left"], t["gray_right"], want=names, sum_mode=SUM_FAST)
        for n in names:
            assert np.array_equal(r[n].cpu().numpy(), first[n], equal_nan=True), n
        del r
    assert (first["left_raw"] == p["disp_gt"]).mean() > 0.9


def test_engine_on_torchs_current_stream_passes_through():
    import torch

    from stereo_match_traditional_b200.api import Engine

    s = torch.cuda.Stream(device="cuda:0")
    with torch.cuda.stream(s):
        e = Engine(0, stream=s.cuda_stream)
        assert e.torch_stream.cuda_stream == s.cuda_stream
        d = torch.arange(40 * 50, device="cuda:0", dtype=torch.float32).reshape(40, 50) % 7
        m = e.median(d, 3)
        s.synchronize()
        ref = e.median(d.cpu().numpy(), 3)
        assert np.array_equal(m.cpu().numpy(), ref)
        e.close()
