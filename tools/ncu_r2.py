"""Driver for `ncu --set full`: one launch (after one warm launch) of each kernel that round 2 works on, at the bench shape
(256 frames = 16 clips x 16 frames of 3x299x299).  python tools/ncu_r2.py [frames] [groups: pool dw stem]"""
import sys
import torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import ops
dev = "cuda"
Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 256
groups = sys.argv[2:] or ["pool", "dw", "stem"]


def st_of(C, Cr):
    st = ops.BNState(C, dev)
    st.scale.zero_(); st.shift.zero_(); st.mean.zero_(); st.rstd.zero_()
    st.scale[:Cr] = torch.rand(Cr, device=dev) + 0.5; st.shift[:Cr] = torch.randn(Cr, device=dev) * 0.1
    st.rstd[:Cr] = 1.0
    st.training = True
    return st


for rep in range(2):
    if "pool" in groups:
        for H, C, Cr in [(147, 128, 128), (74, 256, 256), (37, 768, 728)]:
            y = torch.randn(Fr, H, H, C, device=dev).bfloat16()
            Ho = (H - 1) // 2 + 1
            ys = torch.randn(Fr, Ho, Ho, C, device=dev).bfloat16()
            st, st2 = st_of(C, Cr), st_of(C, Cr)
            out, idx = ops.pool_add_fwd(y, st.scale, st.shift, ys, st2.scale, st2.shift)
            G = torch.randn(Fr, Ho, Ho, C, device=dev).bfloat16()
            gamma = torch.ones(Cr, device=dev); dg = torch.zeros(Cr, device=dev); db = torch.zeros(Cr, device=dev)
            dy = ops.bn_bwd(ops.SRC_POOL, y, st, gamma, dg, db, G=G, idx=idx)
            del y, ys, out, idx, G, dy
    if "dw" in groups:
        for H, C, Cr in [(19, 768, 728), (10, 1536, 1536), (37, 768, 728)]:
            x = torch.randn(Fr, H, H, C, device=dev).bfloat16(); w9 = torch.randn(9, C, device=dev)
            st = st_of(C, Cr)
            dD = torch.randn(Fr, H, H, C, device=dev).bfloat16(); dw9 = torch.zeros(Cr, 1, 3, 3, device=dev); bns = torch.zeros(2, C, device=dev)
            out = ops.dw3x3_fwd(x, w9, st.scale, st.shift, True)
            out2 = ops.dw3x3_fwd(x, w9, None, None, True)
            dz, _ = ops.dw3x3_bwd(dD, x, w9, st.scale, st.shift, True, dw9, bnsum=bns)
            dz2, _ = ops.dw3x3_bwd(dD, x, w9, None, None, True, dw9, add_full=out)
            del x, dD, out, out2, dz, dz2
    if "stem" in groups:
        x = torch.rand(Fr, 3, 299, 299, device=dev)
        w1 = torch.randn(32, 3, 3, 3, device=dev) * 0.3
        y1, p1 = ops.stem_conv1_fwd(x, w1)
        g1 = torch.zeros(32, 3, 3, 3, device=dev)
        ops.stem_conv1_wgrad(x, y1, g1)
        w2 = torch.randn(64, 32, 3, 3, device=dev) * 0.06
        wk, wk_t = ops.pack_conv3x3(w2, True)
        y2, p2 = ops.conv3x3_gemm_fwd(y1, wk)
        dyg = torch.zeros(Fr, 149, 149, 64, device=dev, dtype=torch.bfloat16)
        dyg[:, :147, :147] = y2
        dx = ops.conv3x3_gemm_dgrad(dyg, wk_t)
        gk = torch.zeros(64, 288, device=dev)
        ops.conv3x3_wgrad(dyg, y1, gk)
        del x, y1, y2, dyg, dx
torch.cuda.synchronize()
print("done")
