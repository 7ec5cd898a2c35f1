"""Driver for ncu: the max-pool forward (BN + MaxPool(3,2,1) + skip-BN + add) and the BatchNorm backward through the pool at
block 1's shape (147x147x128) and block 3's (37x37x768 pitch)."""
import sys, torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
dev = "cuda"; Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
for H, C, Cr in [(147, 128, 128), (37, 768, 728)]:
    y = torch.randn(Fr, H, H, C, device=dev).bfloat16()
    Ho = (H - 1) // 2 + 1
    ys = torch.randn(Fr, Ho, Ho, C, device=dev).bfloat16()
    sc = torch.rand(C, device=dev) + 0.5; sh = torch.randn(C, device=dev) * 0.1
    st = ops.BNState(C, dev); st.scale.copy_(sc); st.shift.copy_(sh); st.mean.zero_(); st.rstd.fill_(1.0)
    gamma = torch.ones(Cr, device=dev); dg = torch.zeros(Cr, device=dev); db = torch.zeros(Cr, device=dev)
    for _ in range(2):
        out, idx, ymax = ops.pool_add_fwd(y, sc, sh, ys, sc, sh, True, True)
        out2, _ = ops.pool_add_fwd(y, sc, sh, ys, sc, sh, False, False)
        G = torch.randn_like(out)
        pres = ops.bn_bwd_sums(ymax, G)
        dy = ops.bn_bwd(ops.SRC_POOL, y, st, gamma, dg, db, G=G, idx=idx, presums=pres)
    del y, ys, out, idx, ymax, G, dy, out2
torch.cuda.synchronize()
