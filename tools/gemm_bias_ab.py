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
    res_like = res
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    t0 = timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16, out=out))
    t1 = timeit(lambda: ops.gemm_tn_bias(a, b, bias, True, None, out=out))
    t2 = timeit(lambda: ops.gemm_tn_bias(a, b, bias, False, res, out=out))
    t3 = timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS, out=out))
    if N == 768:
        import os
        res = {}
        dw = torch.zeros(728, 728, device="cuda")
        for tag, env in (("trim", "1"), ("full", "0")):
            os.environ["XCP_GEMM_TRIM"] = env
            res[tag] = (timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16, out=out, n_real=728, k_real=728)),
                        timeit(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS, out=out, n_real=728, k_real=728)),
                        timeit(lambda: ops.gemm_wgrad(a, res_like, dw)))
        os.environ["XCP_GEMM_TRIM"] = "0"
        print("   pad trimming (728 of 768): plain %.1f -> %.1f us | stats %.1f -> %.1f | wgrad %.1f -> %.1f" % (
            res["full"][0], res["trim"][0], res["full"][1], res["trim"][1], res["full"][2], res["trim"][2]), flush=True)
    print("M=%6d N=%4d K=%4d  plain %7.1f us | bias+relu %7.1f | bias+residual %7.1f | stats %7.1f" % (M, N, K, t0, t1, t2, t3), flush=True)
