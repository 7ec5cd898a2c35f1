"""A/B of the row-stream depthwise kernels (csrc/dw_stream.cu) against the kernels they replace (XCP_DW_NO_STREAM=1):
forward and backward, L2 flushed between launches, CUDA events.  python tools/dw_stream_ab.py [frames]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")
PEAK = 6450.6


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


SHAPES = ((19, 768, F), (10, 1024, F), (10, 1536, F), (37, 768, F), (74, 256, F // 2), (147, 128, F // 4))
only = os.environ.get("AB_SHAPES")
for H, C, frames in SHAPES:
    if only and str(H) not in only.split(","):
        continue
    x = torch.randn(frames, H, H, C, device="cuda").to(torch.bfloat16)
    dD = torch.randn(frames, H, H, C, device="cuda").to(torch.bfloat16)
    full = torch.randn(frames, H, H, C, device="cuda").to(torch.bfloat16)
    w9 = torch.randn(9, C, device="cuda")
    sc = torch.rand(C, device="cuda") + 0.5
    sh = torch.randn(C, device="cuda") * 0.1
    out = torch.empty_like(x)
    gb = 2 * x.numel() * 2 / 1e9
    for affine, relu in ((True, True), (False, True)):
        res = {}
        for tag, env in (("old", "1"), ("stream", "0")):
            os.environ["XCP_DW_NO_STREAM"] = env
            a, b = (sc, sh) if affine else (None, None)
            res[tag] = (timeit(lambda: ops.dw3x3_fwd(x, w9, a, b, relu, out=out)), out.clone())
        d = (res["old"][1].float() - res["stream"][1].float()).norm() / res["old"][1].float().norm()
        print("fwd affine=%d relu=%d %3dx%-3dx%-4d F=%-4d  old %7.1f us %.3f | stream %7.1f us %.3f of HBM | rel diff %.1e" % (
            affine, relu, H, H, C, frames, res["old"][0] * 1e3, gb / res["old"][0] * 1e3 / PEAK, res["stream"][0] * 1e3,
            gb / res["stream"][0] * 1e3 / PEAK, d), flush=True)
    if os.environ.get("AB_BWD", "1") != "1":
        continue
    for affine, relu, addf in ((True, True, False), (False, True, True), (True, True, True)):
        res = {}
        gbb = (3 + (1 if addf else 0)) * x.numel() * 2 / 1e9
        for tag, env in (("old", "1"), ("stream", "0")):
            os.environ["XCP_DW_NO_STREAM"] = env
            a, b = (sc, sh) if affine else (None, None)
            gw = torch.zeros(C, 1, 3, 3, device="cuda")

            def run():
                return ops.dw3x3_bwd(dD, x, w9, a, b, relu, gw, add_full=full if addf else None)
            t = timeit(run)
            gw.zero_()
            dz, bns = run()
            res[tag] = (t, dz.clone(), gw.clone(), None if bns is None else bns.clone())
        d = (res["old"][1].float() - res["stream"][1].float()).norm() / res["old"][1].float().norm()
        dw_ = (res["old"][2] - res["stream"][2]).norm() / res["old"][2].norm()
        db = 0.0 if res["old"][3] is None else float((res["old"][3] - res["stream"][3]).norm() / res["old"][3].norm())
        print("bwd affine=%d relu=%d add=%d %3dx%-3dx%-4d F=%-4d  old %7.1f us %.3f | stream %7.1f us %.3f of HBM | dz %.1e dw %.1e bn %.1e" % (
            affine, relu, addf, H, H, C, frames, res["old"][0] * 1e3, gbb / res["old"][0] * 1e3 / PEAK, res["stream"][0] * 1e3,
            gbb / res["stream"][0] * 1e3 / PEAK, d, dw_, db), flush=True)
