"""Per-kernel roofline probe (GPU): times the hot kernels alone at the path's real shapes with CUDA events and
prints achieved GB/s / TFLOP/s against MEASURED_PEAKS.json.  Inputs are larger than L2 (126 MB) or L2 is
flushed between iterations.  Usage: python tools/kernel_bench.py [F] [which...]"""
import json
import math
import os
import sys

import torch

sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops  # noqa: E402

dev = "cuda"
Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
which = set(sys.argv[2:])
peaks = {"hbm_gbs": 6450.6, "bf16_tflops": 1658.4}
if os.path.exists("MEASURED_PEAKS.json"):
    peaks.update(json.load(open("MEASURED_PEAKS.json")))
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def rnd(*s, dtype=torch.bfloat16):
    return torch.randn(*s, device=dev).to(dtype)


def report(name, ms, gbytes=None, tflop=None):
    msg = f"{name:48s} {ms*1e3:9.1f} us"
    if gbytes is not None:
        bw = gbytes / (ms * 1e-3)
        msg += f"  {bw:8.1f} GB/s ({bw / peaks['hbm_gbs'] * 100:5.1f}% of measured HBM)"
    if tflop is not None:
        tf = tflop / (ms * 1e-3)
        msg += f"  {tf:8.1f} TFLOP/s ({tf / peaks['bf16_tflops'] * 100:5.1f}% of measured bf16)"
    print(msg, flush=True)


DW = [(147, 64), (147, 128), (74, 128), (74, 256), (37, 256), (37, 728), (19, 728), (10, 1024), (10, 1536)]
if not which or "dw" in which:
    for H, C in DW:
        x = rnd(Fr, H, H, C); w9 = torch.randn(9, C, device=dev)
        sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
        out = torch.empty_like(x)
        gb = 2 * x.numel() * 2 / 1e9
        report(f"dw_fwd affine+relu  {H}x{H}x{C}", timeit(lambda: ops.dw3x3_fwd(x, w9, sc, sh, True, out=out)), gb)
        report(f"dw_fwd plain        {H}x{H}x{C}", timeit(lambda: ops.dw3x3_fwd(x, w9, None, None, True, out=out)), gb)
        dD = rnd(Fr, H, H, C); dw9 = torch.zeros(C, 1, 3, 3, device=dev); bns = torch.zeros(2, C, device=dev)
        gb3 = 3 * x.numel() * 2 / 1e9
        report(f"dw_bwd affine+relu  {H}x{H}x{C}", timeit(lambda: ops.dw3x3_bwd(dD, x, w9, sc, sh, True, dw9, bnsum=bns)), gb3)
        if H == 19:
            report(f"dw_bwd relu+add_full {H}x{H}x{C}", timeit(lambda: ops.dw3x3_bwd(dD, x, w9, None, None, True, dw9, add_full=dD)), gb3 * 4 / 3)
        del x, out, dD

PW = [(21609, 64, 128), (21609, 128, 128), (5476, 128, 256), (5476, 256, 256), (1369, 256, 728), (1369, 728, 728),
      (361, 728, 728), (361, 728, 1024), (100, 1024, 1536), (100, 1536, 2048)]
if "gemm768" in which:      # alignment probe: 728-wide rows are 1456 B (not 128 B aligned); 768-wide are 1536 B
    PW = [(1369, 728, 728), (1369, 768, 768), (361, 728, 728), (361, 768, 768), (361, 768, 1024), (361, 1024, 1024)]
if not which or "gemm" in which or "gemm768" in which:
    for pix, K, N in PW:
        M = pix * Fr
        a = rnd(M, K); b = rnd(N, K) / math.sqrt(K)
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        tf = 2.0 * M * N * K / 1e12
        gb = (M * K + M * N) * 2 / 1e9
        report(f"gemm fwd+stats  M={M} K={K} N={N}", timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS, out=out)), gb, tf)
        report(f"gemm plain      M={M} K={K} N={N}", timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16, out=out)), gb, tf)
        dw = torch.zeros(N, K, device=dev)
        report(f"gemm wgrad      R={M} P={N} Q={K}", timeit(lambda: ops.gemm_wgrad(out, a, dw)), gb, tf)
        del a, out

if not which or "ew" in which:
    for H, C in [(147, 128), (19, 728)]:
        y = rnd(Fr, H, H, C)
        sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
        gb = 2 * y.numel() * 2 / 1e9
        report(f"bn_act              {H}x{H}x{C}", timeit(lambda: ops.bn_act(y, sc, sh, True)), gb)
        skip = rnd(Fr, H, H, C)
        report(f"bn_add_fwd          {H}x{H}x{C}", timeit(lambda: ops.bn_add_fwd(y, sc, sh, skip)), gb * 1.5)
        Ho = (H - 1) // 2 + 1
        ys = rnd(Fr, Ho, Ho, C)
        report(f"pool_add_fwd        {H}x{H}x{C}", timeit(lambda: ops.pool_add_fwd(y, sc, sh, ys, sc, sh)), (y.numel() + 2 * ys.numel()) * 2 / 1e9)
        st = ops.BNState(C, dev); st.scale.copy_(sc); st.shift.copy_(sh); st.mean.zero_(); st.rstd.fill_(1.0)
        gamma = torch.ones(C, device=dev); dg = torch.zeros(C, device=dev); db = torch.zeros(C, device=dev)
        G = rnd(Fr, H, H, C)
        report(f"bn_bwd direct 2pass {H}x{H}x{C}", timeit(lambda: ops.bn_bwd(ops.SRC_DIRECT, y, st, gamma, dg, db, G=G)), gb * 2.5)
        Gp = rnd(Fr, Ho, Ho, C)
        _, idx = ops.pool_add_fwd(y, sc, sh, ys, sc, sh, want_idx=True)
        report(f"bn_bwd pool 2pass   {H}x{H}x{C}", timeit(lambda: ops.bn_bwd(ops.SRC_POOL, y, st, gamma, dg, db, G=Gp, idx=idx)), gb * 2.5)
        del y, skip, ys, G, Gp, idx

if "stem" in which:
    F_ = Fr
    x1 = rnd(F_, 149, 149, 32); dy2 = rnd(F_, 149, 149, 64)
    gk = torch.zeros(64, 9 * 32, device=dev)
    gb = (x1.numel() + dy2.numel()) * 2 / 1e9
    report("conv2 wgrad fused (9 taps as N tiles)", timeit(lambda: ops.conv3x3_wgrad(dy2, x1, gk)), gb)
    R = F_ * 149 * 149
    def nine():
        for t in range(9):
            sh = (t // 3) * 149 + (t % 3)
            ops._lib.call("xcp_gemm_wgrad", ops._p(dy2), 64, ops._p(x1.view(R, 32)[sh:]), 32,
                          __import__("ctypes").c_void_p(gk.data_ptr() + t * 32 * 4), 9 * 32, R - sh, 64, 32, 0, ops._s())
    report("conv2 wgrad as 9 GEMM launches", timeit(nine), gb)
    xin = torch.rand(F_, 3, 299, 299, device=dev); w1 = torch.randn(32, 3, 3, 3, device=dev)
    report("stem conv1 fwd", timeit(lambda: ops.stem_conv1_fwd(xin, w1)), (xin.numel() * 4 + F_ * 149 * 149 * 32 * 2) / 1e9)
    xu8 = torch.randint(0, 256, (F_, 299, 299, 3), device=dev, dtype=torch.uint8)
    report("stem conv1 fwd (uint8 NHWC frames)", timeit(lambda: ops.stem_conv1_fwd(xu8, w1)), (xu8.numel() + F_ * 149 * 149 * 32 * 2) / 1e9)
    dy1 = rnd(F_, 149, 149, 32); dw1 = torch.zeros(32, 3, 3, 3, device=dev)
    report("stem conv1 wgrad", timeit(lambda: ops.stem_conv1_wgrad(xin, dy1, dw1)), (xin.numel() * 4 + dy1.numel() * 2) / 1e9)
    wk, wkt = ops.pack_conv3x3(torch.randn(64, 32, 3, 3, device=dev), True)
    report("conv2 fwd implicit GEMM + stats", timeit(lambda: ops.conv3x3_gemm_fwd(x1, wk, True)), (x1.numel() + F_ * 147 * 147 * 64) * 2 / 1e9)
    report("conv2 dgrad implicit GEMM", timeit(lambda: ops.conv3x3_gemm_dgrad(dy2, wkt)), (x1.numel() + dy2.numel()) * 2 / 1e9)
