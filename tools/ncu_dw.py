"""ncu driver: depthwise forward (plain ReLU vs fused BN-affine+ReLU) and backward at one shape."""
import sys, torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
dev = "cuda"; Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 32
H = int(sys.argv[2]) if len(sys.argv) > 2 else 147; C = int(sys.argv[3]) if len(sys.argv) > 3 else 128
x = torch.randn(Fr, H, H, C, device=dev).bfloat16(); w9 = torch.randn(9, C, device=dev)
sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
dD = torch.randn(Fr, H, H, C, device=dev).bfloat16(); dw = torch.zeros(C, 1, 3, 3, device=dev); bns = torch.zeros(2, C, device=dev)
for _ in range(2):
    out = ops.dw3x3_fwd(x, w9, None, None, True)
    out = ops.dw3x3_fwd(x, w9, sc, sh, True)
    dz, _ = ops.dw3x3_bwd(dD, x, w9, sc, sh, True, dw, bnsum=bns)
torch.cuda.synchronize()
