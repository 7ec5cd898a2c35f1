"""Production bf16 kernel vs its fp32 twin (csrc/f32_bwd.cu) on the SAME arguments (B200 only).

The fp32 twins are what the plan executor runs in the validation arithmetic, where the whole backward is pinned to the fp32
oracle at 1e-4 per tensor (tests/test_fp32_plan_gpu.py).  Here every production kernel of the backward (and the fused forward
kernels whose saved outputs feed it) is fed bf16-exact inputs and compared with its twin:

  * bf16 outputs must be the correctly rounded fp32 result up to one bf16 ulp (|d| <= 2^-8 |ref| + eps): a wrong term, sign or
    scale in the chain rule shows up at >> 1 ulp, far below the ~1e-2 that bf16 model-level gradient checks can resolve;
  * fp32 outputs (weight gradients, BatchNorm sums, parameter gradients) must agree to 1e-5 relative.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from multimodal_deepfake_detection_b200 import ops  # noqa: E402

DEV = "cuda"
BF, F32 = torch.bfloat16, torch.float32


def bfx(*shape, seed=0, scale=1.0):
    """bf16-exact random tensor, returned as (bf16, fp32) copies of the same values."""
    g = torch.Generator().manual_seed(seed)
    b = (torch.randn(*shape, generator=g) * scale).to(DEV).to(BF)
    return b, b.float()


def one_ulp(out_bf16, ref_f32, extra=0.0):
    """max over elements of |out - ref| / (2^-8 |ref| + floor): <= 1 means 'correctly rounded up to one ulp'."""
    o, r = out_bf16.float(), ref_f32.float()
    floor = r.abs().max() * 2.0 ** -17 + extra
    return ((o - r).abs() / (r.abs() * 2.0 ** -8 + floor)).max().item()


def relmax(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def affine(C, seed, c_real=None):
    g = torch.Generator().manual_seed(seed)
    sc = (torch.rand(C, generator=g) + 0.5) * torch.where(torch.rand(C, generator=g) < 0.15, -1.0, 1.0)
    sh = torch.randn(C, generator=g) * 0.3
    if c_real is not None:
        sc[c_real:] = 0; sh[c_real:] = 0
    return sc.to(DEV), sh.to(DEV)


DW_SHAPES = [(3, 19, 19, 768, 728), (2, 10, 10, 1536, 1536), (2, 37, 37, 256, 256), (1, 74, 74, 128, 128), (2, 8, 8, 768, 728),
             (1, 147, 147, 64, 64), (2, 19, 19, 1024, 1024)]


@pytest.mark.parametrize("shape", DW_SHAPES)
@pytest.mark.parametrize("aff,relu", [(True, True), (False, True), (False, False)])
def test_dw3x3_fwd_twin(shape, aff, relu):
    F_, H, W, C, Cr = shape
    xb, xf = bfx(F_, H, W, C, seed=1)
    g = torch.Generator().manual_seed(2)
    w9 = (torch.randn(9, C, generator=g) * 0.3).to(DEV)
    w9[:, Cr:] = 0
    sc, sh = affine(C, 3, Cr) if aff else (None, None)
    out = ops.dw3x3_fwd(xb, w9, sc, sh, relu)
    ref = ops.dw3x3_fwd(xf, w9, sc, sh, relu)
    assert ref.dtype == F32
    # the production forward folds |scale| into the taps (different fp32 summation order): allow 1 ulp + reassociation noise
    assert one_ulp(out, ref, extra=ref.abs().max().item() * 2e-6) <= 1.01


@pytest.mark.parametrize("shape", DW_SHAPES)
@pytest.mark.parametrize("aff,relu,adds", [(True, True, 0), (False, True, 1), (False, True, 2), (False, False, 2), (False, True, 0)])
def test_dw3x3_bwd_twin(shape, aff, relu, adds):
    """dgrad + ReLU mask + residual adds (identity / stride-2 skip) + 9-tap weight gradient + BN-backward sums."""
    F_, H, W, C, Cr = shape
    xb, xf = bfx(F_, H, W, C, seed=4)
    db, df = bfx(F_, H, W, C, seed=5, scale=0.5)
    g = torch.Generator().manual_seed(6)
    w9 = (torch.randn(9, C, generator=g) * 0.3).to(DEV)
    w9[:, Cr:] = 0
    sc, sh = affine(C, 7, Cr) if aff else (None, None)
    kw_b, kw_f = {}, {}
    if adds == 1:
        kw_b["add_full"], kw_f["add_full"] = bfx(F_, H, W, C, seed=8)
    elif adds == 2:
        kw_b["add_half"], kw_f["add_half"] = bfx(F_, (H + 1) // 2, (W + 1) // 2, C, seed=9)
    dw_b = torch.zeros(Cr, 1, 3, 3, device=DEV); dw_f = torch.zeros(Cr, 1, 3, 3, device=DEV)
    dz_b, sum_b = ops.dw3x3_bwd(db, xb, w9, sc, sh, relu, dw_b, **kw_b)
    dz_f, sum_f = ops.dw3x3_bwd(df, xf, w9, sc, sh, relu, dw_f, **kw_f)
    assert dz_f.dtype == F32
    assert one_ulp(dz_b, dz_f, extra=dz_f.abs().max().item() * 2e-6) <= 1.01
    assert relmax(dw_b, dw_f) < 1e-5
    if aff:
        # the production sums are taken over the UNROUNDED fp32 dz, like the twin's: they agree to summation order
        assert relmax(sum_b[:, :Cr], sum_f[:, :Cr]) < 2e-5


class _St:
    pass


def _bn_state(C, Cr, seed, training):
    g = torch.Generator().manual_seed(seed)
    st = _St()
    gamma = (torch.rand(Cr, generator=g) + 0.5).to(DEV)
    st.mean = torch.zeros(C, device=DEV); st.rstd = torch.zeros(C, device=DEV)
    st.mean[:Cr] = (torch.randn(Cr, generator=g) * 0.2).to(DEV)
    st.rstd[:Cr] = (torch.rand(Cr, generator=g) + 0.7).to(DEV)
    st.scale = torch.zeros(C, device=DEV); st.shift = torch.zeros(C, device=DEV)
    st.scale[:Cr] = gamma * st.rstd[:Cr]
    st.shift[:Cr] = (torch.randn(Cr, generator=g) * 0.2).to(DEV) - st.mean[:Cr] * st.scale[:Cr]
    st.training = training
    return st, gamma


@pytest.mark.parametrize("mode", ["direct", "relu", "pool", "gap", "presums"])
@pytest.mark.parametrize("training", [True, False])
@pytest.mark.parametrize("shape", [(3, 19, 19, 768, 728), (2, 37, 37, 256, 256), (2, 10, 10, 2048, 2048), (1, 21, 17, 64, 64)])
def test_bn_bwd_twin(mode, training, shape):
    """Two-pass BatchNorm backward (batch and frozen statistics) with the ReLU / max-pool / GAP routing folded in."""
    F_, H, W, C, Cr = shape
    yb, yf = bfx(F_, H, W, C, seed=11)
    st, gamma = _bn_state(C, Cr, 12, training)
    kw_b, kw_f = {}, {}
    m = {"direct": ops.SRC_DIRECT, "presums": ops.SRC_DIRECT, "relu": ops.SRC_RELU, "pool": ops.SRC_POOL, "gap": ops.SRC_GAP_RELU}[mode]
    if mode in ("direct", "relu", "presums"):
        kw_b["G"], kw_f["G"] = bfx(F_, H, W, C, seed=13)
    elif mode == "pool":
        Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
        kw_b["G"], kw_f["G"] = bfx(F_, Ho, Wo, C, seed=13)
        idx = torch.randint(0, 9, (F_, Ho, Wo, C), generator=torch.Generator().manual_seed(14), dtype=torch.uint8).to(DEV)
        # taps that fall outside the image never win a real max-pool: point them at the centre tap instead
        oh = torch.arange(Ho, device=DEV).view(1, Ho, 1, 1); ow = torch.arange(Wo, device=DEV).view(1, 1, Wo, 1)
        hi = 2 * oh + (idx // 3).long() - 1; wi = 2 * ow + (idx % 3).long() - 1
        idx = torch.where((hi < 0) | (hi >= H) | (wi < 0) | (wi >= W), torch.full_like(idx, 4), idx).contiguous()
        kw_b["idx"] = kw_f["idx"] = idx
    else:
        kw_b["dfeat"] = kw_f["dfeat"] = (torch.randn(F_, C, generator=torch.Generator().manual_seed(15))).to(DEV)
    if mode == "presums":
        dz = kw_f["G"]
        pres = torch.stack([dz.sum((0, 1, 2)), (dz * yf).sum((0, 1, 2))]).contiguous()
        kw_b["presums"] = kw_f["presums"] = pres
    dg_b, db_b = torch.zeros(Cr, device=DEV), torch.zeros(Cr, device=DEV)
    dg_f, db_f = torch.zeros(Cr, device=DEV), torch.zeros(Cr, device=DEV)
    dy_b = ops.bn_bwd(m, yb, st, gamma, dg_b, db_b, **kw_b)
    dy_f = ops.bn_bwd(m, yf, st, gamma, dg_f, db_f, **kw_f)
    assert dy_f.dtype == F32
    scale_ref = max(dg_f.abs().max().item(), db_f.abs().max().item())
    assert (dg_b - dg_f).abs().max().item() < 2e-5 * scale_ref and (db_b - db_f).abs().max().item() < 2e-5 * scale_ref
    # dy = A*dz + B*y + C: the three terms cancel in train mode, so the comparison floor is the size of the terms
    terms = (dy_f.abs().max().item() + 1.0) * 4e-6
    assert one_ulp(dy_b, dy_f, extra=terms) <= 1.01
    assert (dy_b.float()[..., Cr:] == 0).all() and (dy_f[..., Cr:] == 0).all()


def test_bn_bwd_grid_layout_twin():
    """RELU-source backward scattered onto the zero-padded conv-input grid the stem implicit GEMM consumes."""
    F_, H, W, C = 2, 9, 7, 64
    yb, yf = bfx(F_, H, W, C, seed=21)
    Gb, Gf = bfx(F_, H, W, C, seed=22)
    st, gamma = _bn_state(C, C, 23, True)
    z = torch.zeros(C, device=DEV)
    a = ops.bn_bwd(ops.SRC_RELU, yb, st, gamma, z.clone(), z.clone(), G=Gb, grid_hw=(H + 2, W + 2))
    b = ops.bn_bwd(ops.SRC_RELU, yf, st, gamma, z.clone(), z.clone(), G=Gf, grid_hw=(H + 2, W + 2))
    assert a.shape == b.shape == (F_, H + 2, W + 2, C)
    assert one_ulp(a, b, extra=(b.abs().max().item() + 1.0) * 4e-6) <= 1.01
    assert (a[:, H:].float() == 0).all() and (a[:, :, W:].float() == 0).all()


@pytest.mark.parametrize("shape", [(2, 19, 19, 768, 728), (1, 147, 147, 128, 128), (2, 37, 37, 256, 256), (3, 6, 5, 64, 64)])
@pytest.mark.parametrize("skipbn", [True])
def test_pool_add_and_bn_add_and_gap_twins(shape, skipbn):
    """The fused block tails (their outputs / arg-max taps are what backward consumes)."""
    F_, H, W, C, Cr = shape
    yb, yf = bfx(F_, H, W, C, seed=31)
    sc, sh = affine(C, 32, Cr)
    sc2, sh2 = affine(C, 33, Cr)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    sb, sf = bfx(F_, Ho, Wo, C, seed=34)
    out_b, idx_b = ops.pool_add_fwd(yb, sc, sh, sb, sc2, sh2)
    out_f, idx_f = ops.pool_add_fwd(yf, sc, sh, sf, sc2, sh2)
    assert one_ulp(out_b, out_f, extra=out_f.abs().max().item() * 2e-6) <= 1.01
    # arg-max taps: identical wherever the maximum is unique in fp32 (the bf16 kernel compares raw bf16 values, ties differ)
    # (pad channels have scale = shift = 0: every tap ties there)
    same = (idx_b[..., :Cr] == idx_f[..., :Cr]).float().mean().item()
    assert same > 0.999, same
    # raw winners saved for the backward sums: the same bf16 value in both arithmetics; sums over (ymax, G) == sums over dz, dz*y
    _, _, ym_b = ops.pool_add_fwd(yb, sc, sh, sb, sc2, sh2, want_ymax=True)
    _, _, ym_f = ops.pool_add_fwd(yf, sc, sh, sf, sc2, sh2, want_ymax=True)
    agree = (ym_b.float()[..., :Cr] == ym_f[..., :Cr]).float().mean().item()
    assert agree > 0.999, agree
    Gb, Gf = bfx(F_, Ho, Wo, C, seed=36)
    s_b, s_f = ops.bn_bwd_sums(ym_b, Gb), ops.bn_bwd_sums(ym_f, Gf)
    assert relmax(s_b[:, :Cr], s_f[:, :Cr]) < 1e-3 * (2.0 - agree) and s_b.shape == (2, C)          # (a tie may pick another equal winner)
    st, gamma = _bn_state(C, Cr, 37, True)
    z = lambda: torch.zeros(Cr, device=DEV)  # noqa: E731
    dg1, db1, dg2, db2 = z(), z(), z(), z()
    dy1 = ops.bn_bwd(ops.SRC_POOL, yb, st, gamma, dg1, db1, G=Gb, idx=idx_b)                       # in-kernel routing reduce
    dy2 = ops.bn_bwd(ops.SRC_POOL, yb, st, gamma, dg2, db2, G=Gb, idx=idx_b, presums=ops.bn_bwd_sums(ym_b, Gb))
    ref_scale = max(dg1.abs().max().item(), db1.abs().max().item())
    assert (dg1 - dg2).abs().max().item() < 2e-5 * ref_scale and (db1 - db2).abs().max().item() < 2e-5 * ref_scale
    assert one_ulp(dy2, dy1.float(), extra=(dy1.float().abs().max().item() + 1.0) * 4e-6) <= 1.01
    fb, ff = bfx(F_, H, W, C, seed=35)
    a = ops.bn_add_fwd(yb, sc, sh, fb); b = ops.bn_add_fwd(yf, sc, sh, ff)
    assert one_ulp(a, b, extra=b.abs().max().item() * 2e-6) <= 1.01
    a = ops.bn_add_fwd(yb, sc, sh, fb, sc2, sh2); b = ops.bn_add_fwd(yf, sc, sh, ff, sc2, sh2)
    assert one_ulp(a, b, extra=b.abs().max().item() * 2e-6) <= 1.01
    ga = ops.bn_relu_gap(yb, sc, sh); gb = ops.bn_relu_gap(yf, sc, sh)
    assert relmax(ga, gb) < 1e-5
    aa = ops.bn_act(yb, sc, sh, True); ab = ops.bn_act(yf, sc, sh, True)
    assert one_ulp(aa, ab, extra=ab.abs().max().item() * 2e-6) <= 1.01
    g2a = ops.gather_s2(yb); g2b = ops.gather_s2(yf)
    assert torch.equal(g2a.float(), g2b)


@pytest.mark.parametrize("R,P,Q", [(361 * 5, 728, 728), (21609, 128, 64), (1000, 2048, 1536), (5476 * 2, 256, 128)])
def test_gemm_wgrad_and_dgrad_twins(R, P, Q):
    """tcgen05 weight-gradient (MN-major split-K, fp32 RED epilogue) and data-gradient GEMMs vs the fp32 FFMA twins on
    bf16-exact operands: fp32 outputs to 1e-5, bf16 outputs to one ulp."""
    Pp, Qp = ops.phys(P), ops.phys(Q)
    dyb, dyf = bfx(R, Pp, seed=41, scale=0.5)
    xb, xf = bfx(R, Qp, seed=42)
    dw_b = torch.zeros(P, Q, device=DEV); dw_f = torch.zeros(P, Q, device=DEV)
    ops.gemm_wgrad(dyb, xb, dw_b); ops.gemm_wgrad(dyf, xf, dw_f)
    assert relmax(dw_b, dw_f) < 1e-5
    wb, wf = bfx(Qp, Pp, seed=43, scale=P ** -0.5)           # W^T [K, N] for the data gradient dA = dY . W
    da_b, _ = ops.gemm_tn(dyb, wb, ops.EPI_BF16)
    da_f, _ = ops.gemm_tn(dyf, wf, ops.EPI_BF16)
    assert da_f.dtype == F32
    assert one_ulp(da_b, da_f, extra=da_f.abs().max().item() * 4e-6) <= 1.01
    # forward with the BatchNorm statistics epilogue
    y_b, st_b = ops.gemm_tn(xb, wb.t().contiguous(), ops.EPI_BF16_STATS)
    y_f, st_f = ops.gemm_tn(xf, wf.t().contiguous(), ops.EPI_BF16_STATS)
    assert one_ulp(y_b, y_f, extra=y_f.abs().max().item() * 4e-6) <= 1.01
    assert relmax(st_b.sum(0), st_f.sum(0)) < 2e-5


@pytest.mark.parametrize("F_,Hg", [(2, 37), (1, 75)])
def test_stem_backward_twins(F_, Hg):
    """conv2 implicit-GEMM data / weight gradient and the conv1 weight gradient vs the fp32 direct-convolution twins."""
    g = torch.Generator().manual_seed(51)
    w2 = (torch.randn(64, 32, 3, 3, generator=g) * 0.06).to(DEV).to(BF).float()
    x1b, x1f = bfx(F_, Hg, Hg, 32, seed=52)
    Ho = Hg - 2
    dyb, dyf = bfx(F_, Ho, Ho, 64, seed=53, scale=0.5)
    wk, wk_t = ops.pack_conv3x3(w2, True)
    dy_grid = torch.zeros(F_, Hg, Hg, 64, device=DEV, dtype=BF)
    dy_grid[:, :Ho, :Ho] = dyb
    dx_b = ops.conv3x3_gemm_dgrad(dy_grid, wk_t)
    dx_f = ops.conv3x3_gemm_dgrad(dyf, w2)
    assert one_ulp(dx_b, dx_f, extra=dx_f.abs().max().item() * 4e-6) <= 1.01
    gk = torch.zeros(64, 9 * 32, device=DEV)
    ops.conv3x3_wgrad(dy_grid, x1b, gk)
    gw_b = torch.zeros(64, 32, 3, 3, device=DEV)
    ops.unpack_conv3x3_grad(gk, gw_b)
    gw_f = torch.zeros(64, 32, 3, 3, device=DEV)
    ops.conv3x3_wgrad_f32(x1f, False, dyf, gw_f, 1)
    assert relmax(gw_b, gw_f) < 1e-5
    y_b, st_b = ops.conv3x3_gemm_fwd(x1b, wk)
    y_f, st_f = ops.conv3x3_gemm_fwd(x1f, w2)
    assert one_ulp(y_b, y_f, extra=y_f.abs().max().item() * 4e-6) <= 1.01
    assert relmax(st_b.sum(0), st_f.sum(0)) < 2e-5
    # conv1: fp32 NCHW input, stride 2
    H = 2 * Hg + 1
    x = torch.rand(F_, 3, H, H, generator=g).to(DEV)
    d1b, d1f = bfx(F_, Hg, Hg, 32, seed=54, scale=0.5)
    g1_b = torch.zeros(32, 3, 3, 3, device=DEV); g1_f = torch.zeros(32, 3, 3, 3, device=DEV)
    ops.stem_conv1_wgrad(x, d1b, g1_b); ops.stem_conv1_wgrad(x, d1f, g1_f)
    # the production kernel rounds the im2col of x to bf16 (documented: the stem weight gradient is a bf16 x bf16 GEMM)
    assert relmax(g1_b, g1_f) < 4e-3
    w1 = (torch.randn(32, 3, 3, 3, generator=g) * 0.3).to(DEV)
    y1_b, p_b = ops.stem_conv1_fwd(x, w1)
    y1_f, p_f = ops.stem_conv1_fwd(x, w1, F32)
    # the production kernel is a tf32 tensor-core product (frames and weights rounded to 10 mantissa bits, fp32 accumulate; csrc/
    # stem_conv1.cu): 2^-11 per operand, i.e. ~5e-4 of the output scale instead of fp32's 1e-6
    assert one_ulp(y1_b, y1_f, extra=y1_f.abs().max().item() * 1e-3) <= 1.01
    assert relmax(p_b.sum(0), p_f.sum(0)) < 2e-3
