"""Kernel-level parity (B200 only): every C-ABI kernel against the single torch fp32 op it replaces,
on the same seeded inputs, including the awkward shapes of the path (C=728, H=19/10, M % 128 != 0).
bf16 storage => tolerances are relative to the tensor's max magnitude (written per test)."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from multimodal_deepfake_detection_b200 import _lib, ops  # noqa: E402

DEV = "cuda"


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel_err(a, b):
    a = a.float(); b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def rnd(*shape, seed=0, scale=1.0, dtype=torch.float32):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype)


# ------------------------------------------------------------------------------------------------ GEMMs
@pytest.mark.parametrize("M,N,K", [(256, 128, 64), (1000, 128, 128), (300, 256, 256), (2000, 728, 728),
                                   (1444, 1024, 728), (130, 2048, 1536), (128 * 149 + 5, 64, 128), (77, 8, 8)])
def test_gemm_tn_bf16_stats(M, N, K):
    a = rnd(M, K, seed=1, dtype=torch.bfloat16)
    b = rnd(N, K, seed=2, scale=1.0 / math.sqrt(K), dtype=torch.bfloat16)
    ref = a.float() @ b.float().t()
    out, stats = ops.gemm_tn(a, b, ops.EPI_BF16_STATS)
    assert rel_err(out, ref) < 8e-3          # bf16 rounding of the output only
    s = stats.sum(0)
    assert rel_err(s[0], ref.sum(0)) < 2e-3
    assert rel_err(s[1], (ref * ref).sum(0)) < 2e-3
    cross = ops.gemm_ref(a, b)
    assert rel_err(cross, ref) < 1e-4
    out2, _ = ops.gemm_tn(a, b, ops.EPI_BF16)
    assert torch.equal(out2, out)


@pytest.mark.parametrize("M,N,K,relu,res", [(2000, 768, 768, True, False), (1444, 768, 768, False, True), (300, 128, 64, True, True),
                                             (5000, 256, 128, False, False), (100, 2048, 1536, True, False), (77, 1024, 768, False, True)])
def test_gemm_tn_bias_epilogue(M, N, K, relu, res):
    """inference-plan epilogue: relu?(A B^T + bias + residual) (BatchNorm folded into B, row f-3)"""
    a = rnd(M, K, seed=31, dtype=torch.bfloat16)
    b = rnd(N, K, seed=32, scale=0.05, dtype=torch.bfloat16)
    bias = rnd(N, seed=33, scale=0.5)
    r = rnd(M, N, seed=34, dtype=torch.bfloat16) if res else None
    out = ops.gemm_tn_bias(a, b, bias, relu, r)
    ref = a.float() @ b.float().t() + bias[None, :]
    if res:
        ref = ref + r.float()
    if relu:
        ref = F.relu(ref)
    assert rel_err(out.float(), ref) < 6e-3
    if relu:
        assert float(out.float().min()) >= 0.0


def test_pack_weight_scaled():
    w = rnd(728, 728, seed=35)
    sc = rnd(768, seed=36) * 0.5 + 1.0
    out = ops.pack_weight_scaled(w, sc)
    assert out.shape == (768, 768)
    ref = torch.zeros(768, 768, device=DEV)
    ref[:728, :728] = w * sc[:728, None]
    assert torch.equal(out.float(), ref.bfloat16().float())


@pytest.mark.parametrize("M,epi", [(2000, 0), (2000, 1), (700, 1), (100, 0)])
def test_gemm_tn_pad_trimming(M, epi, monkeypatch):
    """K = N = 768 pitches holding 728 logical channels (zero pad): with n_real / k_real the MMAs skip the padding (46 of 48 K
    steps, 736 of 768 columns, CTA pairs re-split the trimmed N) -- same result as the untrimmed launch, padded output columns
    exactly zero, same BatchNorm partial sums."""
    a = torch.zeros(M, 768, device=DEV, dtype=torch.bfloat16); a[:, :728] = rnd(M, 728, seed=41, dtype=torch.bfloat16)
    b = torch.zeros(768, 768, device=DEV, dtype=torch.bfloat16); b[:728, :728] = rnd(728, 728, seed=42, scale=0.05, dtype=torch.bfloat16)
    monkeypatch.setenv("XCP_GEMM_TRIM", "1")         # off by default (measured slower in the step, csrc/gemm.cu set_trim): A/B hook
    out_t, st_t = ops.gemm_tn(a, b, epi, n_real=728, k_real=728)
    monkeypatch.delenv("XCP_GEMM_TRIM")
    out_f, st_f = ops.gemm_tn(a, b, epi, n_real=728, k_real=728)
    assert torch.equal(out_t, out_f)
    assert float(out_t[:, 728:].abs().max()) == 0.0
    ref = a.float() @ b.float().t()
    assert rel_err(out_t.float(), ref) < 6e-3
    if epi == 1:
        assert rel_err(st_t.sum(0), st_f.sum(0)) < 1e-6
        assert rel_err(st_t.sum(0)[0], ref.sum(0)) < 2e-3
    # weight gradient with the logical (728, 728) shape: the last N tile is issued 224 wide
    dy = a; x = torch.zeros(M, 768, device=DEV, dtype=torch.bfloat16); x[:, :728] = rnd(M, 728, seed=43, dtype=torch.bfloat16)
    dw_t = torch.zeros(728, 728, device=DEV); dw_f = torch.zeros(728, 728, device=DEV)
    monkeypatch.setenv("XCP_GEMM_TRIM", "1")
    ops.gemm_wgrad(dy, x, dw_t)
    monkeypatch.delenv("XCP_GEMM_TRIM")
    ops.gemm_wgrad(dy, x, dw_f)
    assert rel_err(dw_t, dw_f) < 1e-5 and rel_err(dw_t, dy.float()[:, :728].t() @ x.float()[:, :728]) < 1e-4


@pytest.mark.parametrize("M,N,K", [(64, 512, 2048), (300, 2048, 2048), (5, 128, 128)])
def test_gemm_tn_f32_bias(M, N, K):
    a = rnd(M, K, seed=3, dtype=torch.bfloat16)
    b = rnd(N, K, seed=4, scale=1.0 / math.sqrt(K), dtype=torch.bfloat16)
    bias = rnd(N, seed=5)
    ref = a.float() @ b.float().t() + bias
    out, _ = ops.gemm_tn(a, b, ops.EPI_F32, bias=bias)
    assert rel_err(out, ref) < 1e-4


@pytest.mark.parametrize("R,P,Q", [(1000, 128, 64), (5000, 728, 728), (300, 512, 2048), (70000, 128, 128), (50, 1024, 728),
                                   (361 * 16, 2048, 1536)])
def test_gemm_wgrad(R, P, Q):
    dy = rnd(R, P, seed=6, dtype=torch.bfloat16)
    x = rnd(R, Q, seed=7, dtype=torch.bfloat16)
    ref = dy.float().t() @ x.float()
    dw = torch.zeros(P, Q, device=DEV)
    ops.gemm_wgrad(dy, x, dw)
    assert rel_err(dw, ref) < 1e-4
    ops.gemm_wgrad(dy, x, dw)               # accumulates
    assert rel_err(dw, 2 * ref) < 1e-4
    cross = ops.gemm_ref(dy, x, mn_major=True)
    assert rel_err(cross, ref) < 1e-4


@pytest.mark.parametrize("F_,Hg", [(2, 37), (3, 20), (1, 149)])
def test_conv3x3_implicit_gemm(F_, Hg):
    Wg = Hg
    x = rnd(F_, Hg, Wg, 32, seed=8, dtype=torch.bfloat16)
    w = rnd(64, 32, 3, 3, seed=9, scale=0.06)
    wk, wk_t = ops.pack_conv3x3(w)
    wq = wk.float().view(64, 9, 32).permute(0, 2, 1).reshape(64, 32, 3, 3)   # bf16-rounded weights, [O,I,3,3]
    assert rel_err(wq, w) < 8e-3
    x_nchw = x.float().permute(0, 3, 1, 2)
    ref = F.conv2d(x_nchw, wq)                                               # [F,64,Ho,Wo]
    out, stats = ops.conv3x3_gemm_fwd(x, wk)
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 8e-3
    s = stats.sum(0)
    assert rel_err(s[0], ref.sum((0, 2, 3))) < 2e-3
    assert rel_err(s[1], (ref * ref).sum((0, 2, 3))) < 2e-3
    # data gradient: dy on the zero-padded input grid
    dy = rnd(F_, Hg - 2, Wg - 2, 64, seed=10, dtype=torch.bfloat16)
    dy_grid = torch.zeros(F_, Hg, Wg, 64, device=DEV, dtype=torch.bfloat16)
    dy_grid[:, :Hg - 2, :Wg - 2] = dy
    dx = ops.conv3x3_gemm_dgrad(dy_grid, wk_t)
    dx_ref = torch.nn.grad.conv2d_input(x_nchw.shape, wq, dy.float().permute(0, 3, 1, 2))
    assert rel_err(dx.float().permute(0, 3, 1, 2), dx_ref) < 8e-3
    # weight gradient: nine shifted MN-major GEMMs
    gk = torch.zeros(64, 9 * 32, device=DEV)
    ops.conv3x3_wgrad(dy_grid, x, gk)
    gw = torch.zeros(64, 32, 3, 3, device=DEV)
    ops.unpack_conv3x3_grad(gk, gw)
    gw_ref = torch.nn.grad.conv2d_weight(x_nchw, w.shape, dy.float().permute(0, 3, 1, 2))
    assert rel_err(gw, gw_ref) < 1e-4


# ------------------------------------------------------------------------------------------------ stem conv1
@pytest.mark.parametrize("F_,H", [(2, 75), (1, 299), (3, 64)])
def test_stem_conv1(F_, H):
    x = torch.rand(F_, 3, H, H, device=DEV)
    w = rnd(32, 3, 3, 3, seed=11, scale=0.3)
    ref = F.conv2d(x, w, stride=2)
    y, parts = ops.stem_conv1_fwd(x, w)
    assert rel_err(y.float().permute(0, 3, 1, 2), ref) < 8e-3
    s = parts.sum(0)
    # the kernel runs on tf32 tensor-core MMAs (operands rounded to 10 mantissa bits, fp32 accumulate): the statistics are those
    # of the fp32 accumulators of exactly that product -> tight against a tf32-rounded reference, 1e-3 against the fp32 one
    def tf32(t):
        return ((t.contiguous().view(torch.int32) + 0x1000) & ~0x1fff).view(torch.float32)
    ref_t = F.conv2d(tf32(x), tf32(w), stride=2)
    assert rel_err(s[0], ref_t.sum((0, 2, 3))) < 1e-4
    assert rel_err(s[1], (ref_t * ref_t).sum((0, 2, 3))) < 1e-4
    assert rel_err(s[0], ref.sum((0, 2, 3))) < 1e-3 and rel_err(s[1], (ref * ref).sum((0, 2, 3))) < 1e-3
    dy = rnd(*y.shape, seed=12, dtype=torch.bfloat16)
    dw = torch.zeros_like(w)
    ops.stem_conv1_wgrad(x, dy, dw)
    dw_ref = torch.nn.grad.conv2d_weight(x, w.shape, dy.float().permute(0, 3, 1, 2), stride=2)
    # the 27-tap patches are staged in bf16 for the tensor-core reduction (2^-9 per element, random sign): same
    # operand precision as torch's bf16-autocast conv backward; far inside the 1e-2 per-tensor gradient tolerance
    assert rel_err(dw, dw_ref) < 5e-3
    dw_bf = torch.nn.grad.conv2d_weight(x.bfloat16().float(), w.shape, dy.float().permute(0, 3, 1, 2), stride=2)
    assert rel_err(dw, dw_bf) < 1e-4


# ------------------------------------------------------------------------------------------------ depthwise
DW_SHAPES = [(2, 147, 147, 64), (2, 74, 74, 128), (3, 37, 37, 256), (2, 19, 19, 728), (3, 10, 10, 1024), (1, 5, 7, 16), (9, 37, 37, 728),
             (2, 2, 2, 728), (1, 33, 31, 8),
             # tiny square maps of the audio model (register-resident forward kernel), incl. the 768-pitch width
             (7, 4, 4, 768), (5, 8, 8, 256), (3, 2, 2, 1536), (4, 3, 3, 64), (2, 1, 1, 64), (130, 4, 4, 768), (3, 5, 5, 1024), (2, 7, 7, 8),
             (2, 6, 6, 128), (2, 8, 4, 64),
             # row-walking kernels (W = 10 / 15 / 19) on non-square maps: every H mod 3, H < 3, generic and compile-time pitch
             (2, 11, 10, 1024), (1, 5, 19, 768), (2, 2, 15, 64), (1, 1, 10, 64), (2, 12, 10, 1536), (1, 20, 15, 256), (2, 3, 19, 768)]


@pytest.mark.parametrize("shape", DW_SHAPES)
@pytest.mark.parametrize("affine,relu", [(False, False), (False, True), (True, True), (True, False)])
def test_dw3x3_fwd_bwd(shape, affine, relu):
    F_, H, W, C = shape
    x = rnd(F_, H, W, C, seed=13, dtype=torch.bfloat16)
    w = rnd(C, 1, 3, 3, seed=14, scale=0.4)
    w9 = ops.pack_dw(w)
    scale = (rnd(C, seed=15) * 0.5 + 1.0) if affine else None      # includes some negative / small scales
    shift = rnd(C, seed=16, scale=0.3) if affine else None
    if affine:      # corner cases of the folded BN affine: negative scale, zero scale with +/- shift, zero-padded channel (0, 0)
        scale[0], scale[1], shift[1], scale[2], shift[2], scale[3], shift[3] = -0.7, 0.0, 0.3, 0.0, -0.3, 0.0, 0.0
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wt = w.clone().requires_grad_(True)
    a = xt
    if affine:
        a = a * scale[None, :, None, None] + shift[None, :, None, None]
    if relu:
        a = F.relu(a)
    ref = F.conv2d(a, wt, padding=1, groups=C)
    out = ops.dw3x3_fwd(x, w9, scale, shift, relu)
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 8e-3
    # backward
    dD = rnd(F_, H, W, C, seed=17, dtype=torch.bfloat16)
    ref.backward(dD.float().permute(0, 3, 1, 2))
    gw = torch.zeros_like(w)
    dz, bnsum = ops.dw3x3_bwd(dD, x, w9, scale, shift, relu, gw)
    assert rel_err(gw, wt.grad) < 2e-3
    # dz is the gradient wrt the pre-activation z (= scale*x+shift, or x): compare through the chain rule
    dz_ref = xt.grad
    if affine:
        sc = scale[None, :, None, None]
        dz_f = dz.float().permute(0, 3, 1, 2)
        assert rel_err(dz_f * sc, dz_ref) < 1e-2
        zsum = dz_f.sum((0, 2, 3))
        zysum = (dz_f * xt.detach()).sum((0, 2, 3))
        # the kernel sums the un-rounded fp32 dz; the check sums the bf16-rounded dz -> loose tolerance
        assert rel_err(bnsum[0], zsum) < 2e-2
        assert rel_err(bnsum[1], zysum) < 2e-2
    else:
        assert rel_err(dz.float().permute(0, 3, 1, 2), dz_ref) < 1e-2


def test_dw3x3_bwd_residual_adds():
    F_, H, W, C = 2, 19, 19, 728
    x = rnd(F_, H, W, C, seed=18, dtype=torch.bfloat16)
    w9 = ops.pack_dw(rnd(C, 1, 3, 3, seed=19, scale=0.4))
    dD = rnd(F_, H, W, C, seed=20, dtype=torch.bfloat16)
    full = rnd(F_, H, W, C, seed=21, dtype=torch.bfloat16)
    half = rnd(F_, 10, 10, C, seed=22, dtype=torch.bfloat16)
    gw = torch.zeros(C, 1, 3, 3, device=DEV)
    base, _ = ops.dw3x3_bwd(dD, x, w9, None, None, True, gw)
    both, _ = ops.dw3x3_bwd(dD, x, w9, None, None, True, gw, add_full=full, add_half=half)
    ref = base.float() + full.float()
    ref[:, ::2, ::2] += half.float()
    assert rel_err(both, ref) < 8e-3


def test_dw3x3_bwd_residual_adds_inside_the_relu_mask():
    """relu = 2 (unmaterialised block input, executor.block_forward(inp_st=...)): the residual gradients are gradients wrt the
    ACTIVATED input relu(scale*x + shift), so they go inside the mask: dz = mask * (conv_transpose(dD) + add_half@even)."""
    F_, H, W, C = 2, 21, 21, 64
    x = rnd(F_, H, W, C, seed=23, dtype=torch.bfloat16)
    w9 = ops.pack_dw(rnd(C, 1, 3, 3, seed=24, scale=0.4))
    scale, shift = rnd(C, seed=25) * 0.5 + 1.0, rnd(C, seed=26, scale=0.3)
    dD = rnd(F_, H, W, C, seed=27, dtype=torch.bfloat16)
    half = rnd(F_, 11, 11, C, seed=28, dtype=torch.bfloat16)
    gw = torch.zeros(C, 1, 3, 3, device=DEV)
    # reference through autograd: a = relu(scale*x+shift) feeds the depthwise conv AND (at even pixels) a second consumer
    xt = x.float().permute(0, 3, 1, 2)
    z = (xt * scale[None, :, None, None] + shift[None, :, None, None]).requires_grad_(True)
    a = F.relu(z)
    wt = rnd(C, 1, 3, 3, seed=24, scale=0.4)
    out = F.conv2d(a, wt, padding=1, groups=C)
    loss = (out * dD.float().permute(0, 3, 1, 2)).sum() + (a[:, :, ::2, ::2] * half.float().permute(0, 3, 1, 2)).sum()
    loss.backward()
    dz, bns = ops.dw3x3_bwd(dD, x, w9, scale, shift, 2, gw, add_half=half)
    assert rel_err(dz.float().permute(0, 3, 1, 2), z.grad) < 8e-3
    assert rel_err(bns[0], z.grad.sum((0, 2, 3))) < 2e-2
    assert rel_err(bns[1], (z.grad * xt).sum((0, 2, 3))) < 2e-2


# ------------------------------------------------------------------------------------------------ BN & elementwise
@pytest.mark.parametrize("shape", [(4, 19, 19, 728), (2, 37, 37, 256), (3, 10, 10, 2048), (1, 149, 149, 32)])
def test_bn_finalize_and_apply(shape):
    F_, H, W, C = shape
    y = rnd(F_, H, W, C, seed=23, dtype=torch.bfloat16) * 2 + 0.5
    yf = y.float().permute(0, 3, 1, 2).contiguous()
    gamma = rnd(C, seed=24) * 0.2 + 1
    beta = rnd(C, seed=25, scale=0.2)
    rm = rnd(C, seed=26, scale=0.1)
    rv = torch.rand(C, device=DEV) + 0.5
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(gamma); bn.bias.copy_(beta); bn.running_mean.copy_(rm); bn.running_var.copy_(rv)
    bn.train()
    ref = bn(yf)
    # partials as a GEMM epilogue could produce them: per 128-row tile sums
    y2 = y.float().view(-1, C)
    M = y2.shape[0]
    mt = (M + 127) // 128
    pad = torch.zeros(mt * 128, C, device=DEV); pad[:M] = y2
    parts = torch.stack([pad.view(mt, 128, C).sum(1), (pad * pad).view(mt, 128, C).sum(1)], 1).contiguous()
    rm2, rv2 = rm.clone(), rv.clone()
    st = ops.bn_finalize(parts, M, gamma, beta, rm2, rv2, True)
    assert rel_err(rm2, bn.running_mean) < 1e-5 and rel_err(rv2, bn.running_var) < 1e-5
    out = ops.bn_act(y, st.scale, st.shift, False)
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 8e-3
    out = ops.bn_act(y, st.scale, st.shift, True)
    assert rel_err(out.float().permute(0, 3, 1, 2), F.relu(ref)) < 8e-3
    bn.eval()
    st_e = ops.bn_finalize(None, M, gamma, beta, bn.running_mean, bn.running_var, False)
    out = ops.bn_act(y, st_e.scale, st_e.shift, False)
    assert rel_err(out.float().permute(0, 3, 1, 2), bn(yf)) < 8e-3
    # GAP
    feat = ops.bn_relu_gap(y, st.scale, st.shift)
    assert rel_err(feat, F.relu(ref).mean((2, 3))) < 1e-3
    # gather
    g = ops.gather_s2(y, st.scale, st.shift, True)
    assert rel_err(g.float().permute(0, 3, 1, 2), F.relu(ref)[:, :, ::2, ::2]) < 8e-3
    g = ops.gather_s2(y)
    assert torch.equal(g, y[:, ::2, ::2].contiguous())
    # bn + add
    skip = rnd(F_, H, W, C, seed=27, dtype=torch.bfloat16)
    o = ops.bn_add_fwd(y, st.scale, st.shift, skip)
    assert rel_err(o.float().permute(0, 3, 1, 2), ref + skip.float().permute(0, 3, 1, 2)) < 8e-3


@pytest.mark.parametrize("shape", [(2, 19, 19, 728), (2, 37, 37, 256), (1, 147, 147, 128), (3, 4, 4, 64), (2, 5, 3, 16)])
def test_pool_add_fwd_and_bn_bwd_pool(shape):
    F_, H, W, C = shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    y = rnd(F_, H, W, C, seed=28, dtype=torch.bfloat16)
    ys = rnd(F_, Ho, Wo, C, seed=29, dtype=torch.bfloat16)
    gamma = rnd(C, seed=30) * 0.5 + 1
    gamma[::7] *= -1                                 # negative BN scales must pool correctly
    beta = rnd(C, seed=31, scale=0.2)
    gamma_s = rnd(C, seed=32) * 0.2 + 1
    beta_s = rnd(C, seed=33, scale=0.2)

    def stats_parts(t):
        t2 = t.float().view(-1, C)
        return torch.stack([t2.sum(0), (t2 * t2).sum(0)], 0)[None].contiguous(), t2.shape[0]

    p, n = stats_parts(y); st = ops.bn_finalize(p, n, gamma, beta, None, None, True)
    p, n = stats_parts(ys); st_s = ops.bn_finalize(p, n, gamma_s, beta_s, None, None, True)
    out, idx = ops.pool_add_fwd(y, st.scale, st.shift, ys, st_s.scale, st_s.shift)

    yt = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    yst = ys.float().permute(0, 3, 1, 2).requires_grad_(True)
    g_t = gamma.clone().requires_grad_(True); b_t = beta.clone().requires_grad_(True)
    z = F.batch_norm(yt, None, None, g_t, b_t, True, 0.1, 1e-5)
    zs = F.batch_norm(yst, None, None, gamma_s, beta_s, True, 0.1, 1e-5)
    ref = F.max_pool2d(z, 3, 2, 1) + zs
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 8e-3
    G = rnd(F_, Ho, Wo, C, seed=34, dtype=torch.bfloat16)
    ref.backward(G.float().permute(0, 3, 1, 2))
    dgamma = torch.zeros(C, device=DEV); dbeta = torch.zeros(C, device=DEV)
    dy = ops.bn_bwd(ops.SRC_POOL, y, st, gamma, dgamma, dbeta, G=G, idx=idx)
    assert rel_err(dy.float().permute(0, 3, 1, 2), yt.grad) < 2e-2
    assert rel_err(dgamma, g_t.grad) < 1e-2 and rel_err(dbeta, b_t.grad) < 1e-2
    # skip branch: direct mode
    dg2 = torch.zeros(C, device=DEV); db2 = torch.zeros(C, device=DEV)
    dys = ops.bn_bwd(ops.SRC_DIRECT, ys, st_s, gamma_s, dg2, db2, G=G)
    assert rel_err(dys.float().permute(0, 3, 1, 2), yst.grad) < 2e-2


@pytest.mark.parametrize("mode", ["relu", "gap", "eval_direct"])
def test_bn_bwd_modes(mode):
    F_, H, W, C = 3, 10, 10, 2048
    y = rnd(F_, H, W, C, seed=35, dtype=torch.bfloat16)
    gamma = rnd(C, seed=36) * 0.2 + 1
    beta = rnd(C, seed=37, scale=0.2)
    rm = rnd(C, seed=38, scale=0.1); rv = torch.rand(C, device=DEV) + 0.5
    y2 = y.float().view(-1, C)
    parts = torch.stack([y2.sum(0), (y2 * y2).sum(0)], 0)[None].contiguous()
    training = mode != "eval_direct"
    st = ops.bn_finalize(parts if training else None, y2.shape[0], gamma, beta, rm.clone(), rv.clone(), training)
    yt = y.float().permute(0, 3, 1, 2).requires_grad_(True)
    g_t = gamma.clone().requires_grad_(True); b_t = beta.clone().requires_grad_(True)
    z = F.batch_norm(yt, rm.clone(), rv.clone(), g_t, b_t, training, 0.1, 1e-5)
    dgamma = torch.zeros(C, device=DEV); dbeta = torch.zeros(C, device=DEV)
    if mode == "relu":
        G = rnd(F_, H, W, C, seed=39, dtype=torch.bfloat16)
        F.relu(z).backward(G.float().permute(0, 3, 1, 2))
        dy = ops.bn_bwd(ops.SRC_RELU, y, st, gamma, dgamma, dbeta, G=G)
    elif mode == "gap":
        dfeat = rnd(F_, C, seed=40)
        F.relu(z).mean((2, 3)).backward(dfeat)
        dy = ops.bn_bwd(ops.SRC_GAP_RELU, y, st, gamma, dgamma, dbeta, dfeat=dfeat)
    else:
        G = rnd(F_, H, W, C, seed=41, dtype=torch.bfloat16)
        z.backward(G.float().permute(0, 3, 1, 2))
        dy = ops.bn_bwd(ops.SRC_DIRECT, y, st, gamma, dgamma, dbeta, G=G)
    assert rel_err(dy.float().permute(0, 3, 1, 2), yt.grad) < 2e-2
    assert rel_err(dgamma, g_t.grad) < 1e-2 and rel_err(dbeta, b_t.grad) < 1e-2


def test_layout_and_packing():
    x = rnd(3, 5, 11, 13, seed=42)
    n = ops.nchw_to_nhwc(x)
    assert torch.equal(n, x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    back = ops.nhwc_to_nchw(n)
    assert torch.equal(back, n.float().permute(0, 3, 1, 2).contiguous())
    w = rnd(728, 256, seed=43)
    wb, wt = ops.pack_weight(w)
    assert torch.equal(wb, w.to(torch.bfloat16)) and torch.equal(wt, w.t().contiguous().to(torch.bfloat16))
    a = rnd(6, 3, 13, 1, seed=44, scale=20.0)
    up = ops.bilinear_up(a, 64)
    ref = F.interpolate(a, size=(64, 64), mode="bilinear", align_corners=False)
    assert rel_err(up, ref) < 1e-5
    assert torch.equal(ops.cast_bf16(w), w.to(torch.bfloat16))


# ------------------------------------------------------------------------------------------------ LSTM / head
@pytest.mark.parametrize("B,T,H", [(4, 16, 128), (2, 5, 32), (3, 7, 512), (8, 120, 512), (11, 9, 512), (1, 1, 512), (5, 6, 256),
                                   (16, 33, 256), (16, 16, 128), (8, 120, 128), (9, 3, 128), (1, 1, 128)])
def test_lstm_fwd_bwd(B, T, H):
    lstm = torch.nn.LSTM(2048, H, 1, batch_first=True).to(DEV)
    with torch.no_grad():   # the recurrent weights are held in bf16 by the kernel
        lstm.weight_hh_l0.copy_(lstm.weight_hh_l0.to(torch.bfloat16).float())
        lstm.weight_ih_l0.copy_(lstm.weight_ih_l0.to(torch.bfloat16).float())
    x = rnd(B, T, 2048, seed=45, scale=0.5).to(torch.bfloat16).float().requires_grad_(True)
    ref, (hn_ref, cn_ref) = lstm(x)
    w_ih_b, _ = ops.pack_weight(lstm.weight_ih_l0.detach(), want_t=False)
    w_hh_b, w_hh_t = ops.pack_weight(lstm.weight_hh_l0.detach())
    xproj, _ = ops.gemm_tn(x.detach().view(B * T, 2048).to(torch.bfloat16), w_ih_b, ops.EPI_F32)
    h, gates, cst, hn, cn = ops.lstm_fwd(xproj, lstm.bias_ih_l0.detach(), lstm.bias_hh_l0.detach(), w_hh_t, B, T, H)
    assert rel_err(h, ref) < 2e-4 and rel_err(hn, hn_ref[0]) < 2e-4 and rel_err(cn, cn_ref[0]) < 2e-4
    dout = rnd(B, T, H, seed=46)
    ref.backward(dout)
    dbi = torch.zeros(4 * H, device=DEV); dbh = torch.zeros(4 * H, device=DEV)
    dgates, hprev = ops.lstm_bwd(dout, None, None, gates, cst, h, w_hh_b, dbi, dbh, B, T, H)
    assert rel_err(dbi, lstm.bias_ih_l0.grad) < 1e-2
    dwih = torch.zeros(4 * H, 2048, device=DEV)
    ops.gemm_wgrad(dgates, x.detach().view(B * T, 2048).to(torch.bfloat16), dwih)
    assert rel_err(dwih, lstm.weight_ih_l0.grad) < 1e-2
    dwhh = torch.zeros(4 * H, H, device=DEV)
    ops.gemm_wgrad(dgates, hprev, dwhh)
    assert rel_err(dwhh, lstm.weight_hh_l0.grad) < 1e-2
    _, w_ih_t = ops.pack_weight(lstm.weight_ih_l0.detach())
    dx, _ = ops.gemm_tn(dgates, w_ih_t, ops.EPI_F32)
    assert rel_err(dx.view(B, T, 2048), x.grad) < 1e-2


def test_lstm_cluster_kernels_match_single_cta_kernels(monkeypatch):
    """H = 512: the cluster kernels (W_hh in registers, DSMEM h exchange) against the single-CTA kernels on the same inputs,
    incl. the initial dh_n / dc_n gradients, and their speed at the audio model's T = 120."""
    import time
    B, T, H = 8, 120, 512
    g = torch.Generator().manual_seed(80)
    w = (torch.randn(4 * H, H, generator=g) * 0.04).to(DEV)
    w_b, w_t = ops.pack_weight(w)
    xproj = (torch.randn(B * T, 4 * H, generator=g) * 0.5).to(DEV)
    bi, bh = (torch.randn(4 * H, generator=g) * 0.1).to(DEV), (torch.randn(4 * H, generator=g) * 0.1).to(DEV)
    dout = torch.randn(B, T, H, generator=g).to(DEV)
    dhn, dcn = torch.randn(B, H, generator=g).to(DEV), torch.randn(B, H, generator=g).to(DEV)

    def run():
        torch.cuda.synchronize(); t0 = time.perf_counter()
        h, gates, cst, hn, cn = ops.lstm_fwd(xproj, bi, bh, w_t, B, T, H)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        dbi = torch.zeros(4 * H, device=DEV); dbh = torch.zeros(4 * H, device=DEV)
        dgates, hprev = ops.lstm_bwd(dout, dhn, dcn, gates, cst, h, w_b, dbi, dbh, B, T, H)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        return (h, gates, cst, hn, cn, dgates.float(), hprev.float(), dbi, dbh), (t1 - t0) * 1e3, (t2 - t1) * 1e3
    run()
    new, f_new, b_new = run()
    monkeypatch.setenv("XCP_LSTM_NO_CLUSTER", "1")
    run()
    old, f_old, b_old = run()
    monkeypatch.delenv("XCP_LSTM_NO_CLUSTER")
    print("lstm H=512 T=120 B=8: fwd %.2f -> %.2f ms, bwd %.2f -> %.2f ms" % (f_old, f_new, b_old, b_new))
    names = ("h", "gates", "c", "hn", "cn", "dgates", "hprev", "dbi", "dbh")
    for n, a, b in zip(names, new, old):
        tol = 2e-2 if n in ("dgates", "dbi", "dbh") else 2e-3
        assert rel_err(a, b) < tol, (n, rel_err(a, b))
    assert f_new < f_old and b_new < b_old


def test_head_linear_bce():
    B, K, N = 4, 128, 1024
    a = rnd(B, K, seed=47).requires_grad_(True)
    W = rnd(N, K, seed=48, scale=0.1).requires_grad_(True)
    b = rnd(N, seed=49, scale=0.1).requires_grad_(True)
    mask = (torch.rand(B, N, device=DEV) > 0.3).to(torch.uint8)
    ds = 1.0 / 0.7
    ref = F.relu(F.linear(a, W, b)) * mask.float() * ds
    out = ops.linear_small_fwd(a.detach(), W.detach(), b.detach(), 1, mask, ds)
    assert rel_err(out, ref) < 1e-5
    delta = rnd(B, N, seed=50)
    ref.backward(delta)
    dW = torch.zeros_like(W); db = torch.zeros_like(b)
    din = ops.linear_small_bwd(delta, out, ds, a.detach(), W.detach(), dW, db)
    assert rel_err(dW, W.grad) < 1e-4 and rel_err(db, b.grad) < 1e-4 and rel_err(din, a.grad) < 1e-4
    z = rnd(B, 1, seed=51).requires_grad_(True)
    y = torch.tensor([[1.0], [0.0], [1.0], [0.0]], device=DEV)
    loss_ref = F.binary_cross_entropy(torch.sigmoid(z), y)
    loss_ref.backward()
    probs, loss, dz = ops.bce_fwd_bwd(z.detach(), y)
    assert rel_err(probs, torch.sigmoid(z)) < 1e-5 and abs(loss.item() - loss_ref.item()) < 1e-5 and rel_err(dz, z.grad) < 1e-4
    z2 = z.detach().clone().requires_grad_(True)
    l2 = F.binary_cross_entropy_with_logits(z2, y * 0.9 + 0.05)
    l2.backward()
    _, loss_s, dz_s = ops.bce_fwd_bwd(z2.detach(), y, smoothing=0.1)
    assert abs(loss_s.item() - l2.item()) < 1e-5 and rel_err(dz_s, z2.grad) < 1e-4


def test_pack_multi_matches_per_tensor_packs():
    """executor.PackCache.prefetch (one xcp_pack_multi launch) == the per-tensor pack kernels, incl. 728 -> 768 zero padding,
    in-place refresh after a parameter update and a table that changes."""
    from multimodal_deepfake_detection_b200 import executor as ex
    g = torch.Generator().manual_seed(70)
    pws = [torch.randn(n, k, 1, 1, generator=g).to(DEV) for n, k in ((128, 64), (728, 256), (728, 728), (1024, 728), (2048, 1536), (40, 24))]
    dws = [torch.randn(c, 1, 3, 3, generator=g).to(DEV) for c in (64, 728, 1536, 24)]
    items = [(w, "pw") for w in pws] + [(w, "dw") for w in dws]
    cache = ex.PackCache()
    cache.prefetch(items)

    def check():
        for w in pws:
            a, at = cache.pw(w)
            b, bt = ops.pack_weight(w.view(w.shape[0], w.shape[1]), True, pad=True)
            assert torch.equal(a, b) and torch.equal(at, bt)
        for w in dws:
            assert torch.equal(cache.dw(w), ops.pack_dw(w, pad=True))
    n0 = _lib.launch_count()
    check()
    ptrs = [cache.pw(w)[0].data_ptr() for w in pws]
    for w in pws + dws:
        w.mul_(1.5)                                   # bumps Tensor._version: every pack is stale again
    cache.prefetch(items)
    check()
    assert ptrs == [cache.pw(w)[0].data_ptr() for w in pws]          # refreshed in place
    ex.bump_param_epoch()                             # raw-pointer update (FusedAdam): stale through the epoch
    pws[2].data.add_(1.0)
    cache.prefetch(items[:5] + items[6:])             # a different table
    check()


def test_bce_loss_module_matches_nn_bceloss():
    """BCELoss (train_audio.py:20,39) vs torch.nn.BCELoss incl. saturated probabilities (log clamp at -100, gradient
    denominator clamp at 1e-12), a non-unit upstream gradient and a (B,1) / (B,) mix of shapes."""
    from multimodal_deepfake_detection_b200 import BCELoss
    g = torch.Generator().manual_seed(60)
    for n in (1, 4, 32, 300):
        p = torch.rand(n, 1, generator=g).to(DEV)
        if n >= 4:
            p[0], p[1], p[2] = 0.0, 1.0, 1e-30
        y = torch.randint(0, 2, (n, 1), generator=g).float().to(DEV)
        pr = p.clone().requires_grad_(True)
        po = p.clone().requires_grad_(True)
        lr = torch.nn.BCELoss()(pr, y)
        lo = BCELoss()(po, y)
        (lr * 3.0).backward(); (lo * 3.0).backward()
        assert abs(lr.item() - lo.item()) <= 1e-5 * max(1.0, abs(lr.item()))
        assert rel_err(po.grad, pr.grad) < 1e-5
    with torch.no_grad():
        assert BCELoss()(p, y).requires_grad is False
    with pytest.raises(_lib.XcpError):
        BCELoss()(p, y[:-1])


def test_arcface_and_fusion_against_golden(golden):
    import numpy as np

    def t(k):
        return torch.from_numpy(np.asarray(golden[k])).to(DEV)

    lab = t("D_labels")
    # visual ArcFace (s=30, m=0.5) + CE  (train_visual.py:455-474,532)
    dw = torch.zeros(2, 32, device=DEV)
    logits, loss, dx = ops.arcface_loss(t("D2_emb"), t("D2_w"), lab, 30.0, 0.5, 0, dw=dw)
    assert rel_err(logits, t("D2_logits")) < 1e-4 and abs(loss.item() - float(golden["D2_loss"])) < 1e-4
    assert rel_err(dx, t("D2_grad_emb")) < 1e-3 and rel_err(dw, t("D2_grad_w")) < 1e-3
    lg, _, _ = ops.arcface_loss(t("D2_emb"), t("D2_w"), None, 30.0, 0.5)
    assert rel_err(lg, t("D2_logits_nolabel")) < 1e-4
    # fusion region (train_au_face.py:659-674), dropout off
    v, a = t("D_v_tok"), t("D_a_tok")
    pooled, loss_reg, dv, da = ops.fusion_pool_reg(v, a, 0.2, 0.1)
    W0, b0, W3, b3 = t("D_embed::0.weight"), t("D_embed::0.bias"), t("D_embed::3.weight"), t("D_embed::3.bias")
    h = ops.linear_small_fwd(pooled, W0, b0, 1)
    e = ops.linear_small_fwd(h, W3, b3, 0)
    darc = torch.zeros(2, 128, device=DEV)
    logits, loss_cls, de = ops.arcface_loss(e, t("D_arc_w"), lab, 30.0, 0.30, 1, class_w=t("D_class_weights"), gamma=2.0, dw=darc)
    assert rel_err(logits, t("D_logits")) < 1e-4
    assert abs((loss_cls + loss_reg).item() - float(golden["D_loss"])) < 1e-4
    assert rel_err(darc, t("D_grad_arc_w")) < 1e-3
    dW3 = torch.zeros_like(W3); db3 = torch.zeros_like(b3)
    dh = ops.linear_small_bwd(de, None, 1.0, h, W3, dW3, db3)
    dW0 = torch.zeros_like(W0); db0 = torch.zeros_like(b0)
    dpooled = ops.linear_small_bwd(dh, h, 1.0, pooled, W0, dW0, db0)
    assert rel_err(dW0, t("D_grad_embed0_w")) < 1e-3
    ops.fusion_pool_bwd(dpooled, dv, da)
    assert rel_err(dv, t("D_grad_v_tok")) < 1e-3 and rel_err(da, t("D_grad_a_tok")) < 1e-3


def test_adam_and_clip():
    n = 100003
    p = rnd(n, seed=52); g = rnd(n, seed=53, scale=0.01)
    pt = p.clone().requires_grad_(True)
    opt = torch.optim.Adam([pt], lr=1e-3, weight_decay=1e-4)
    m = torch.zeros(n, device=DEV); v = torch.zeros(n, device=DEV)
    ss = torch.zeros((), device=DEV)
    for step in range(1, 4):
        pt.grad = g.clone()
        torch.nn.utils.clip_grad_norm_([pt], 1.0)
        opt.step()
        ops.grad_sumsq(g, ss)
        ops.adam_step(p, g, m, v, 1e-3, 0.9, 0.999, 1e-8, 1e-4, False, step, sumsq=ss, max_norm=1.0)
        assert rel_err(p, pt.detach()) < 1e-5


# ------------------------------------------------------------------------------------------------ uint8 ingest
def test_stem_conv1_uint8_nhwc_ingest_matches_float_path():
    """Raw uint8 NHWC frames (video_dataloader.py:27-35 on-disk layout) through the stem == the reference's
    float(T,3,H,W)/255 tensor through the stem (the in-kernel u8/255 and torch's scalar division differ by <= 1 ulp of the
    fp32 input, i.e. at most a bf16 rounding flip in the output; 4x fewer input bytes)."""
    g = torch.Generator().manual_seed(31)
    u8 = torch.randint(0, 256, (3, 75, 61, 3), generator=g, dtype=torch.uint8).to(DEV)
    xf = (u8.permute(0, 3, 1, 2).float() / 255.0).contiguous()
    w = rnd(32, 3, 3, 3, seed=32, scale=0.3)
    y_f, p_f = ops.stem_conv1_fwd(xf, w)
    y_u, p_u = ops.stem_conv1_fwd(u8, w)
    assert rel_err(y_u, y_f) < 1e-3 and rel_err(p_u.sum(0), p_f.sum(0)) < 1e-5
    ref = F.conv2d(xf, w, stride=2)
    assert rel_err(y_u.float().permute(0, 3, 1, 2), ref) < 4e-3
    # BN partial sums come from the fp32 accumulators of the tf32 tensor-core product (operands rounded to 10 mantissa bits)
    assert rel_err(p_u[:, 0].sum(0), ref.sum((0, 2, 3))) < 1e-3
    assert rel_err(p_u[:, 1].sum(0), (ref * ref).sum((0, 2, 3))) < 1e-3
    dy = rnd(*y_f.shape, seed=33, dtype=torch.bfloat16)
    dw_f = torch.zeros_like(w); dw_u = torch.zeros_like(w)
    ops.stem_conv1_wgrad(xf, dy, dw_f)
    ops.stem_conv1_wgrad(u8, dy, dw_u)
    assert rel_err(dw_u, dw_f) < 1e-5                                   # same patches; split-K RED order differs


def test_model_accepts_uint8_clips():
    from multimodal_deepfake_detection_b200 import XceptionLSTMV
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMV(32).to(DEV).eval()
    g = torch.Generator().manual_seed(34)
    u8 = torch.randint(0, 256, (2, 3, 75, 75, 3), generator=g, dtype=torch.uint8).to(DEV)            # (B,T,H,W,3)
    xf = (u8.permute(0, 1, 4, 2, 3).float() / 255.0).contiguous()                                     # (B,T,3,H,W)
    with torch.no_grad():
        a = m.extract_features(u8)
        b = m.extract_features(xf)
    assert a.shape == (2, 3, 2048) and rel_err(a, b) < 5e-3


@pytest.mark.parametrize("shape", [(2, 147, 147, 128), (3, 37, 37, 768), (2, 19, 19, 1024), (1, 6, 5, 16), (2, 8, 8, 64), (1, 3, 3, 8)])
def test_pool_add_fwd_interior_fast_path_ties_and_no_index_variant(shape):
    """All-positive BatchNorm scales take the interior fast path (no border select, no sign flip).  Inputs drawn from seven
    values so that nearly every window has ties: the arg-max must be the FIRST maximum in row-major window order, exactly
    like F.max_pool2d(return_indices=True) (Xception.py:86 MaxPool2d(3, 2, 1)); y[arg-max] is the raw winner; the inference
    variant without index bookkeeping writes bit-identical sums."""
    F_, H, W, C = shape
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    g = torch.Generator().manual_seed(H * 131 + C)
    y = (torch.randint(-3, 4, shape, generator=g).float() * 0.5).to(DEV).bfloat16()
    ys = rnd(F_, Ho, Wo, C, seed=61, dtype=torch.bfloat16)
    sc = (torch.rand(C, generator=g) + 0.5).to(DEV); sh = (torch.randn(C, generator=g) * 0.2).to(DEV)
    scs = (torch.rand(C, generator=g) + 0.5).to(DEV); shs = (torch.randn(C, generator=g) * 0.2).to(DEV)
    out, idx, ymax = ops.pool_add_fwd(y, sc, sh, ys, scs, shs, True, True)
    out_noidx, none_idx = ops.pool_add_fwd(y, sc, sh, ys, scs, shs, False, False)
    assert none_idx is None and torch.equal(out, out_noidx)
    z = y.float().permute(0, 3, 1, 2) * sc[None, :, None, None] + sh[None, :, None, None]
    ref, ind = F.max_pool2d(z, 3, 2, 1, return_indices=True)
    zs = ys.float().permute(0, 3, 1, 2) * scs[None, :, None, None] + shs[None, :, None, None]
    assert rel_err(out.float().permute(0, 3, 1, 2), ref + zs) < 8e-3
    ho = torch.arange(Ho, device=DEV)[None, None, :, None]; wo = torch.arange(Wo, device=DEV)[None, None, None, :]
    tap = (ind // W - (2 * ho - 1)) * 3 + (ind % W - (2 * wo - 1))
    assert torch.equal(idx.permute(0, 3, 1, 2).long(), tap)
    y_at = torch.gather(y.float().permute(0, 3, 1, 2).reshape(F_, C, H * W), 2, ind.reshape(F_, C, Ho * Wo)).reshape(F_, C, Ho, Wo)
    assert torch.equal(ymax.float().permute(0, 3, 1, 2), y_at)


@pytest.mark.parametrize("F_,H", [(2, 299), (3, 64), (1, 75)])
def test_stem_conv1_affine_inference_form(F_, H):
    """Inference stem (row f-3): relu(bn1(conv1(x))) of Xception.py:168-170 in ONE pass -- eval-mode bn1 scale folded into the
    filter rows, shift + ReLU in the epilogue -- against torch and against the two-pass path (conv1, then BN + ReLU)."""
    x = torch.rand(F_, 3, H, H, device=DEV)
    w = rnd(32, 3, 3, 3, seed=11, scale=0.3)
    scale = rnd(32, seed=12) * 0.3 + 1.0
    scale[::5] *= -1
    shift = rnd(32, seed=13, scale=0.3)
    out = ops.stem_conv1_fwd_affine(x, w, scale, shift)
    ref = F.relu(F.conv2d(x, w, stride=2) * scale[None, :, None, None] + shift[None, :, None, None])
    assert rel_err(out.float().permute(0, 3, 1, 2), ref) < 8e-3
    y, _ = ops.stem_conv1_fwd(x, w)
    two_pass = ops.bn_act(y, scale, shift, True)
    assert rel_err(out, two_pass) < 8e-3          # the two-pass path rounds the raw conv output to bf16 before the affine
    u8 = (x * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous()
    out_u8 = ops.stem_conv1_fwd_affine(u8, w, scale, shift)
    ref_u8 = F.relu(F.conv2d(u8.permute(0, 3, 1, 2).float() / 255.0, w, stride=2) * scale[None, :, None, None] + shift[None, :, None, None])
    assert rel_err(out_u8.float().permute(0, 3, 1, 2), ref_u8) < 8e-3


@pytest.mark.parametrize("F_,S,C", [(5, 4, 768), (1100, 4, 768), (3, 2, 1536), (700, 2, 1024), (9, 4, 64), (5, 8, 768), (300, 8, 256)])
@pytest.mark.parametrize("mode", ["plain", "relu", "affine_relu", "full", "half", "both", "pre_half"])
def test_dw3x3_small_bwd_tiny_maps(F_, S, C, mode):
    """The register-resident depthwise backward of the audio model's 8x8 / 4x4 / 2x2 maps (csrc/dw_small_bwd.cu) in every fused mode
    against torch autograd: dgrad, 9-tap weight gradient (logical channels of a padded pitch), BN-affine / ReLU mask,
    identity-skip and stride-2-skip gradient adds (outside and inside the mask), BatchNorm-backward sums; enough frames for
    several images per thread and a ragged last frame group."""
    Cr = 728 if C == 768 else C
    x = rnd(F_, S, S, C, seed=71, dtype=torch.bfloat16)
    wt = rnd(Cr, 1, 3, 3, seed=72, scale=0.4)
    dD = rnd(F_, S, S, C, seed=73, dtype=torch.bfloat16)
    if Cr < C:
        x[..., Cr:] = 0; dD[..., Cr:] = 0
    w9 = ops.pack_dw(wt, pad=True)                  # [9][768]: zero taps in the pad channels
    affine = mode in ("affine_relu", "pre_half")
    relu = mode != "plain"
    scale = shift = None
    if affine:
        scale = torch.zeros(C, device=DEV); shift = torch.zeros(C, device=DEV)
        scale[:Cr] = rnd(Cr, seed=74) * 0.5 + 1.0; shift[:Cr] = rnd(Cr, seed=75, scale=0.3)
    Sh = (S + 1) // 2
    full = rnd(F_, S, S, C, seed=76, dtype=torch.bfloat16) if mode in ("full", "both") else None
    half = rnd(F_, Sh, Sh, C, seed=77, dtype=torch.bfloat16) if mode in ("half", "both", "pre_half") else None
    gw = torch.zeros(Cr, 1, 3, 3, device=DEV)
    dz, bns = ops.dw3x3_bwd(dD, x, w9, scale, shift, 2 if mode == "pre_half" else relu, gw, add_full=full, add_half=half)

    xt = x.float().permute(0, 3, 1, 2)[:, :Cr]
    z = xt.clone()
    if affine:
        z = z * scale[None, :Cr, None, None] + shift[None, :Cr, None, None]
    z.requires_grad_(True)
    wr = wt.clone().requires_grad_(True)
    a = F.relu(z) if relu else z
    out = F.conv2d(a, wr, padding=1, groups=Cr)
    loss = (out * dD.float().permute(0, 3, 1, 2)[:, :Cr]).sum()
    if mode == "pre_half":              # the stride-2 consumer reads the ACTIVATED input: its gradient goes inside the mask
        loss = loss + (a[:, :, ::2, ::2] * half.float().permute(0, 3, 1, 2)[:, :Cr]).sum()
    loss.backward()
    ref = z.grad.clone()
    if full is not None:
        ref = ref + full.float().permute(0, 3, 1, 2)[:, :Cr]
    if half is not None and mode != "pre_half":
        ref[:, :, ::2, ::2] += half.float().permute(0, 3, 1, 2)[:, :Cr]
    got = dz.float().permute(0, 3, 1, 2)
    assert rel_err(got[:, :Cr], ref) < 8e-3
    if Cr < C:
        assert float(got[:, Cr:].abs().max()) == 0.0 or full is not None or half is not None
    assert rel_err(gw, wr.grad) < 2e-3
    if affine:
        assert rel_err(bns[0][:Cr], ref.sum((0, 2, 3))) < 2e-2
        assert rel_err(bns[1][:Cr], (ref * xt).sum((0, 2, 3))) < 2e-2
