"""Does the weight-gradient GEMM co-run with the BatchNorm backward?  Serial vs two-stream timing at the middle-flow shape."""
import sys
import torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops

dev = "cuda"
F_, H, C, Cr = 256, 19, 768, 728
M = F_ * H * H
y = torch.randn(F_, H, H, C, device=dev).bfloat16()
G = torch.randn(F_, H, H, C, device=dev).bfloat16()
d = torch.randn(M, C, device=dev).bfloat16()
dy = torch.randn(M, C, device=dev).bfloat16()
dw = torch.zeros(Cr, Cr, device=dev)
st = ops.BNState(C, dev); st.scale.fill_(1.0); st.shift.zero_(); st.mean.zero_(); st.rstd.fill_(1.0); st.training = True
gamma = torch.ones(Cr, device=dev); dg = torch.zeros(Cr, device=dev); db = torch.zeros(Cr, device=dev)
pres = torch.randn(2, C, device=dev)
side = torch.cuda.Stream()
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)


def bn():
    return ops.bn_bwd(ops.SRC_DIRECT, y, st, gamma, dg, db, G=G, presums=pres)


def wg():
    ops.gemm_wgrad(dy, d, dw)


def both():
    ev = torch.cuda.Event(); ev.record()
    side.wait_event(ev)
    with torch.cuda.stream(side):
        wg()
    bn()
    torch.cuda.current_stream().wait_stream(side)


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2] * 1e3


print("bn apply alone %.1f us | wgrad alone %.1f us | serial %.1f us | two streams %.1f us" % (
    timeit(bn), timeit(wg), timeit(lambda: (wg(), bn())), timeit(both)))
