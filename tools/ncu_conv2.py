"""Driver for ncu: the stem conv2 implicit GEMM (forward + statistics, dgrad) at the bench shape (256 frames of 149x149x32)."""
import sys, torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
dev = "cuda"; F_ = int(sys.argv[1]) if len(sys.argv) > 1 else 256
x1 = torch.randn(F_, 149, 149, 32, device=dev).bfloat16(); dy2 = torch.randn(F_, 149, 149, 64, device=dev).bfloat16()
wk, wkt = ops.pack_conv3x3(torch.randn(64, 32, 3, 3, device=dev), True)
for _ in range(3):
    ops.conv3x3_gemm_fwd(x1, wk, True)
    ops.conv3x3_gemm_dgrad(dy2, wkt)
torch.cuda.synchronize()
