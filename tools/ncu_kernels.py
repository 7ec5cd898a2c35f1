"""Tiny driver for ncu: launches each hot kernel a few times at the middle-flow shape (F frames of 19x19x728, stored
with the 768-channel physical pitch) and the largest entry-flow shape (147x147x128)."""
import sys, math, torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
dev = "cuda"; Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for H, C, Cr in [(19, 768, 728), (147, 128, 128)]:
    F_ = Fr if H == 19 else max(Fr // 4, 1)
    x = torch.randn(F_, H, H, C, device=dev).bfloat16(); w9 = torch.randn(9, C, device=dev)
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
    dD = torch.randn(F_, H, H, C, device=dev).bfloat16(); dw9 = torch.zeros(Cr, 1, 3, 3, device=dev); bns = torch.zeros(2, C, device=dev)
    a = x.view(-1, C); b = (torch.randn(C, C, device=dev) / math.sqrt(C)).bfloat16()
    dw = torch.zeros(Cr, Cr, device=dev)
    st = ops.BNState(C, dev); st.scale.copy_(sc); st.shift.copy_(sh); st.mean.zero_(); st.rstd.fill_(1.0)
    gamma = torch.ones(Cr, device=dev); dg = torch.zeros(Cr, device=dev); db = torch.zeros(Cr, device=dev)
    for _ in range(2):
        out = ops.dw3x3_fwd(x, w9, sc, sh, True)
        dz, _ = ops.dw3x3_bwd(dD, x, w9, sc, sh, True, dw9, bnsum=bns)
        y, stt = ops.gemm_tn(a, b, ops.EPI_BF16_STATS)
        ops.gemm_wgrad(y, a, dw)
        ops.bn_bwd(ops.SRC_DIRECT, x, st, gamma, dg, db, G=dD)
    del x, dD, a, y, out, dz
torch.cuda.synchronize()
