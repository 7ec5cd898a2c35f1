"""A/B of the depthwise forward kernels on the narrow / tiny maps: TMA halo-tile kernel (XCP_DW_NO_ROWS=1 XCP_DW_NO_SMALL=1)
vs the register-window kernels.  python tools/dw_rows_ab.py [frames]"""
import os
import sys

import torch

sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops  # noqa: E402

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda")


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


for H, C, frames in ((19, 768, F), (10, 1024, F), (10, 1536, F), (15, 256, 960), (8, 768, 960), (4, 768, 960), (2, 1536, 960)):
    x = torch.randn(frames, H, H, C, device="cuda").to(torch.bfloat16)
    w9 = torch.randn(9, C, device="cuda")
    sc = torch.rand(C, device="cuda") + 0.5
    sh = torch.randn(C, device="cuda") * 0.1
    out = torch.empty_like(x)
    gb = 2 * x.numel() * 2 / 1e9
    res = {}
    for tag, env, envc in (("tma", "1", "1"), ("rows", "0", "1"), ("regs", "0", "0")):      # regs = best register-window kernel
        os.environ["XCP_DW_NO_ROWS"] = env
        os.environ["XCP_DW_NO_SMALL"] = env
        os.environ["XCP_DW_NO_ROWSC"] = envc
        res[tag] = (timeit(lambda: ops.dw3x3_fwd(x, w9, sc, sh, True, out=out)), out.clone())
    d = (res["tma"][1].float() - res["regs"][1].float()).norm() / res["tma"][1].float().norm()
    print("dw_fwd affine+relu %3dx%-3dx%-4d F=%-4d  tma %7.1f us %6.0f GB/s | generic rows %7.1f us | regs %7.1f us %6.0f GB/s | rel diff %.1e" % (
        H, H, C, frames, res["tma"][0] * 1e3, gb / res["tma"][0] * 1e3, res["rows"][0] * 1e3, res["regs"][0] * 1e3, gb / res["regs"][0] * 1e3, d))
