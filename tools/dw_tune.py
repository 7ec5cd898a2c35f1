"""Offline geometry sweep for the depthwise kernels: for each Xception shape try (CTAs/SM, strips, row slices, tile rows)
candidates through the XCP_DW_GEOM / XCP_DW_MINB hooks (re-read on every call) and print the fastest."""
import os, sys
import torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
SHAPES = [(147, 128), (74, 256), (37, 256), (37, 728), (19, 728), (10, 1536)]
dev = "cuda"; Fr = 128
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    try:
        fn(); fn()
        ts = []
        for _ in range(4):
            flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        return sorted(ts)[1] * 1e3
    except Exception:
        return -1.0


for H, C in SHAPES:
    x = torch.randn(Fr, H, H, C, device=dev).bfloat16(); w9 = torch.randn(9, C, device=dev)
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
    dD = torch.randn(Fr, H, H, C, device=dev).bfloat16(); dw = torch.zeros(C, 1, 3, 3, device=dev); bns = torch.zeros(2, C, device=dev)
    out = torch.empty_like(x)
    fns = {"fwd": lambda: ops.dw3x3_fwd(x, w9, sc, sh, True, out=out), "bwd": lambda: ops.dw3x3_bwd(dD, x, w9, sc, sh, True, dw, bnsum=bns)}
    for which in ("fwd", "bwd"):
        os.environ.pop("XCP_DW_GEOM", None); os.environ.pop("XCP_DW_MINB", None)
        base = timeit(fns[which])
        res = []
        for minb in (1, 2):
            maxw = 15 if minb == 1 else 7
            halo = (800 if which == "fwd" else 470) if minb == 1 else (400 if which == "fwd" else 200)
            cands = set()
            for ns in range(1, maxw + 1):
                if ns > 1 and 4 * (ns - 1) >= H: break
                for rs in range(1, maxw // ns + 1):
                    th_max = halo // (4 * ns + 2) - 2
                    if th_max < 1: continue
                    n0 = (H + th_max - 1) // th_max
                    for n_h in range(n0, n0 + 2):
                        th = (H + n_h - 1) // n_h
                        if rs > th or ns * rs < maxw * 0.6: continue
                        cands.add((ns, rs, th))
            for ns, rs, th in sorted(cands):
                os.environ["XCP_DW_MINB"] = str(minb); os.environ["XCP_DW_GEOM"] = "%d,%d,%d" % (ns, rs, th)
                t = timeit(fns[which])
                if t > 0: res.append((t, minb, ns, rs, th))
        res.sort()
        print(which, H, C, "default %.1fus | best:" % base, ["%.1fus minb=%d ns=%d rs=%d th=%d" % r for r in res[:4]], "n=%d" % len(res), flush=True)
    del x, dD, out
