"""A/B on one box: the classifier head as one launch per direction (csrc/head_fused.cu) against the five-launch path
(xcp_linear_small_fwd/bwd + sigmoid + BCE), both replayed from CUDA graphs so launch gaps are the graph's, not Python's.
usage: python tools/head_ab.py [B] [H]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_deepfake_detection_b200 import ops  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H = int(sys.argv[2]) if len(sys.argv) > 2 else 128
dev = torch.device("cuda:0")
T, Wd = 16, 1024
g = torch.Generator().manual_seed(0)
x = torch.randn(B, T, H, generator=g).to(dev)
wb = []
for n, k in [(Wd, H), (Wd, Wd), (Wd, Wd), (Wd, Wd), (1, Wd)]:
    wb += [(torch.randn(n, k, generator=g) * (2.0 / k) ** 0.5).to(dev), torch.zeros(n, device=dev)]
dwb = [torch.zeros_like(t) for t in wb]
y = torch.randint(0, 2, (B, 1), generator=g).float().to(dev)
masks = (torch.rand(4, B, Wd, generator=g) >= 0.3).to(torch.uint8).to(dev)


def fused():
    acts, z, prob, loss, dz = ops.head_mlp_fwd(x, None, wb, 0.3, None, 1, y)
    return acts, prob, loss, dz


def fused_bwd(acts, dz):
    return ops.head_mlp_bwd(dz, None, None, x, None, acts, 1 / 0.7, wb, dwb)


def unfused():
    a = [x[:, -1, :].contiguous()]
    for li in range(4):
        a.append(ops.linear_small_fwd(a[-1], wb[2 * li], wb[2 * li + 1], 1, masks[li], 1 / 0.7))
    z = ops.linear_small_fwd(a[-1], wb[8], wb[9], 0)
    p = ops.sigmoid_fwd(z)
    loss, dp = ops.bce_prob_fwd_bwd(p, y)
    return a, p, loss, dp


def unfused_bwd(a, p, dp):
    dz = ops.sigmoid_bwd(p, dp)
    d = ops.linear_small_bwd(dz, None, 1.0, a[4], wb[8], dwb[8], dwb[9])
    for li in range(3, -1, -1):
        d = ops.linear_small_bwd(d, a[li + 1], 1 / 0.7, a[li], wb[2 * li], dwb[2 * li], dwb[2 * li + 1])
    return d


def timed(fn, reps=50):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        fn()
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); graph.replay(); e1.record()
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1e3


acts, prob, loss, dz = fused()
a, p, l2, dp = unfused()
print("B=%d H=%d  forward (+loss): fused %.1f us | five launches %.1f us" % (B, H, timed(fused), timed(unfused)))
print("            backward:        fused %.1f us | six launches  %.1f us" % (timed(lambda: fused_bwd(acts, dz)), timed(lambda: unfused_bwd(a, p, dp))))
