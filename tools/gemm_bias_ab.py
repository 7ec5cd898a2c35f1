"""A/B: pointwise GEMM with the plain bf16 epilogue vs the folded-BatchNorm epilogue (bias + ReLU / residual)."""
import sys
import torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops

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
    return sorted(ts)[len(ts) // 2] * 1e3


for M, N, K in ((92416, 768, 768), (256 * 147 * 147 // 4, 128, 128), (256 * 74 * 74 // 2, 256, 256), (25600, 1536, 1024), (25600, 2048, 1536)):
    a = torch.randn(M, K, device="cuda").bfloat16()
    b = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    bias = torch.randn(N, device="cuda")
    res = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t0 = timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16, out=out))
    t1 = timeit(lambda: ops.gemm_tn_bias(a, b, bias, True, None, out=out))
    t2 = timeit(lambda: ops.gemm_tn_bias(a, b, bias, False, res, out=out))
    t3 = timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS, out=out))
    print("M=%6d N=%4d K=%4d  plain %7.1f us | bias+relu %7.1f | bias+residual %7.1f | stats %7.1f" % (M, N, K, t0, t1, t2, t3), flush=True)
