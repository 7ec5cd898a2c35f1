"""Memory-safety and race evidence without compute-sanitizer (B200 only).

`compute-sanitizer` is closed on this GPU pool (profiles/r2_compute_sanitizer_closed_on_pool.txt), so the memcheck / racecheck
tier of SURVEY.md §4 is covered by bounds checks of our own, at the odd shapes of the path (C = 728 -> 768 pitch, H = 19 / 10 / 2,
M % 128 != 0, partial tiles):

  * every INPUT lives inside a larger allocation whose guard bands (64 KB on either side) hold NaN / 0xFF: an out-of-bounds
    READ that reaches the arithmetic poisons the result;
  * every OUTPUT the wrappers allocate (ops.py's torch.empty / empty_like / zeros are intercepted) lives between guard bands
    filled with a canary byte: an out-of-bounds WRITE breaks the canary;
  * every kernel runs three times: outputs without atomically accumulated sums must be bit-identical run to run (a shared-memory
    or mbarrier race shows up as run-to-run noise), accumulated ones within fp32 summation-order noise; nothing may be NaN.
"""
import contextlib

import pytest
import torch

pytestmark = pytest.mark.gpu

from multimodal_deepfake_detection_b200 import ops  # noqa: E402

DEV = "cuda"
GUARD = 64 * 1024          # bytes on either side
CANARY = 0xA5
BF, F32 = torch.bfloat16, torch.float32


class _Guarded:
    """torch-namespace proxy for ops.py: allocations come out of canary-filled buffers and are remembered for checking."""

    def __init__(self):
        self.bufs = []

    def _alloc(self, shape, dtype, device, fill=None):
        shape = tuple(int(s) for s in (shape if isinstance(shape, (tuple, list, torch.Size)) else (shape,)))
        n = 1
        for s in shape:
            n *= s
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        pad = (-nbytes) % 256
        raw = torch.full((GUARD + nbytes + pad + GUARD,), CANARY, dtype=torch.uint8, device=device)
        body = raw[GUARD:GUARD + nbytes].view(dtype).view(shape)
        if fill is not None:
            body.fill_(fill)
        self.bufs.append((raw, nbytes))
        return body

    def empty(self, *size, dtype=None, device=None, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        return self._alloc(shape, dtype or F32, device or DEV)

    def zeros(self, *size, dtype=None, device=None, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        return self._alloc(shape, dtype or F32, device or DEV, fill=0)

    def empty_like(self, t, **kw):
        return self._alloc(t.shape, t.dtype, t.device)

    def zeros_like(self, t, **kw):
        return self._alloc(t.shape, t.dtype, t.device, fill=0)

    def __getattr__(self, name):
        return getattr(torch, name)

    def check(self):
        torch.cuda.synchronize()
        for raw, nbytes in self.bufs:
            assert bool((raw[:GUARD] == CANARY).all()), "out-of-bounds write BELOW an output buffer"
            assert bool((raw[GUARD + nbytes:] == CANARY).all()), "out-of-bounds write ABOVE an output buffer"
        self.bufs = []


@contextlib.contextmanager
def guarded_outputs():
    g = _Guarded()
    real = ops.torch
    ops.torch = g
    try:
        yield g
    finally:
        ops.torch = real


def gin(t: torch.Tensor) -> torch.Tensor:
    """Copy an input into the middle of a NaN / 0xFF-filled allocation."""
    t = t.contiguous()
    nbytes = t.numel() * t.element_size()
    pad = (-nbytes) % 256
    raw = torch.full((GUARD + nbytes + pad + GUARD,), 0xFF, dtype=torch.uint8, device=t.device)
    body = raw[GUARD:GUARD + nbytes].view(t.dtype).view(t.shape)
    body.copy_(t)
    body._guard_owner = raw
    return body


def rnd(*shape, seed=0, scale=1.0, dtype=BF):
    g = torch.Generator().manual_seed(seed)
    return gin((torch.randn(*shape, generator=g) * scale).to(DEV).to(dtype))


def _flat(res):
    if torch.is_tensor(res):
        return [res]
    if isinstance(res, (tuple, list)):
        out = []
        for r in res:
            out += _flat(r)
        return out
    return []


def run3(fn, exact=True, tol=2e-5):
    """fn() -> tensor(s).  Three guarded runs: canaries intact, no NaN, run-to-run identical (or within `tol` of the max)."""
    outs = []
    for _ in range(3):
        with guarded_outputs() as g:
            res = [t.clone() for t in _flat(fn())]
            g.check()
        outs.append(res)
    for res in outs:
        for t in res:
            if t.dtype.is_floating_point:
                assert bool(torch.isfinite(t.float()).all()), "NaN / Inf in an output: an out-of-bounds read reached the arithmetic"
    for a, b in zip(outs[0], outs[1]):
        if exact or not a.dtype.is_floating_point:
            assert torch.equal(a, b), "run-to-run difference in a deterministic output"
        else:
            assert (a.float() - b.float()).abs().max().item() <= tol * (a.float().abs().max().item() + 1e-30)
    for a, b in zip(outs[0], outs[2]):
        if exact or not a.dtype.is_floating_point:
            assert torch.equal(a, b)
    return outs[0]


@pytest.mark.parametrize("M,N,K", [(361 * 3 + 5, 768, 768), (77, 8, 8), (130, 2048, 1536), (128 * 3 + 1, 128, 64)])
def test_gemms(M, N, K):
    a, b = rnd(M, K, seed=1), rnd(N, K, seed=2, scale=K ** -0.5)
    run3(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS)[0])
    run3(lambda: ops.gemm_tn(a, b, ops.EPI_BF16_STATS)[1], exact=False)          # per-CTA partials: CTA -> tile assignment is static
    run3(lambda: ops.gemm_tn(a, b, ops.EPI_F32)[0])
    dy = rnd(M, N, seed=3)

    def wgrad():
        dw = ops.torch.zeros((N, K), device=DEV, dtype=F32)
        ops.gemm_wgrad(dy, a, dw)
        return dw
    run3(wgrad, exact=False)


@pytest.mark.parametrize("F_,Hg", [(2, 21), (1, 38)])
def test_stem(F_, Hg):
    g = torch.Generator().manual_seed(4)
    x = gin(torch.rand(F_, 3, 2 * Hg + 1, 2 * Hg + 1, generator=g).to(DEV))
    w1 = gin((torch.randn(32, 3, 3, 3, generator=g) * 0.3).to(DEV))
    y1 = run3(lambda: ops.stem_conv1_fwd(x, w1)[0])[0]
    y1 = gin(y1)
    w2 = (torch.randn(64, 32, 3, 3, generator=g) * 0.06).to(DEV)
    wk, wk_t = ops.pack_conv3x3(w2, True)
    wk, wk_t = gin(wk), gin(wk_t)
    run3(lambda: ops.conv3x3_gemm_fwd(y1, wk)[0])
    dyg = torch.zeros(F_, Hg, Hg, 64, device=DEV, dtype=BF)
    dyg[:, :Hg - 2, :Hg - 2] = torch.randn(F_, Hg - 2, Hg - 2, 64, generator=g).to(DEV).to(BF)
    dyg = gin(dyg)
    run3(lambda: ops.conv3x3_gemm_dgrad(dyg, wk_t))

    def wg():
        gk = ops.torch.zeros((64, 288), device=DEV, dtype=F32)
        ops.conv3x3_wgrad(dyg, y1, gk)
        return gk
    run3(wg, exact=False)

    def wg1():
        g1 = ops.torch.zeros((32, 3, 3, 3), device=DEV, dtype=F32)
        ops.stem_conv1_wgrad(x, y1, g1)
        return g1
    run3(wg1, exact=False)


DW = [(3, 19, 19, 768, 728), (2, 10, 10, 1536, 1536), (2, 37, 37, 256, 256), (1, 74, 74, 128, 128), (2, 8, 8, 768, 728),
      (3, 2, 2, 1536, 1536), (2, 15, 15, 256, 256), (1, 147, 147, 64, 64), (2, 5, 3, 64, 64)]


@pytest.mark.parametrize("shape", DW)
def test_depthwise(shape):
    F_, H, W, C, Cr = shape
    x, dD = rnd(F_, H, W, C, seed=5), rnd(F_, H, W, C, seed=6, scale=0.5)
    w9 = torch.randn(9, C, device=DEV) * 0.3
    w9[:, Cr:] = 0
    w9 = gin(w9)
    sc = torch.rand(C, device=DEV) + 0.5; sh = torch.randn(C, device=DEV) * 0.2
    sc[Cr:] = 0; sh[Cr:] = 0
    sc, sh = gin(sc), gin(sh)
    run3(lambda: ops.dw3x3_fwd(x, w9, sc, sh, True))
    run3(lambda: ops.dw3x3_fwd(x, w9, None, None, False))
    addf = rnd(F_, H, W, C, seed=7)
    addh = rnd(F_, (H + 1) // 2, (W + 1) // 2, C, seed=8)

    def bwd(aff, **kw):
        def f():
            dw = ops.torch.zeros((Cr, 1, 3, 3), device=DEV, dtype=F32)
            dz, bns = ops.dw3x3_bwd(dD, x, w9, sc if aff else None, sh if aff else None, True, dw, **kw)
            return dz
        return f
    run3(bwd(True))                       # dz is deterministic; dw / bnsum are RED sums (checked for parity elsewhere)
    run3(bwd(False, add_full=addf))
    run3(bwd(False, add_half=addh))


@pytest.mark.parametrize("shape", [(2, 19, 19, 768, 728), (1, 37, 37, 256, 256), (3, 6, 5, 64, 64), (2, 10, 10, 2048, 2048)])
def test_elementwise_and_bn_backward(shape):
    F_, H, W, C, Cr = shape
    y = rnd(F_, H, W, C, seed=9)
    st = ops.BNState(C, DEV)
    st.scale.zero_(); st.shift.zero_(); st.mean.zero_(); st.rstd.zero_()
    st.scale[:Cr] = torch.rand(Cr, device=DEV) + 0.5; st.shift[:Cr] = torch.randn(Cr, device=DEV) * 0.2
    st.mean[:Cr] = torch.randn(Cr, device=DEV) * 0.1; st.rstd[:Cr] = torch.rand(Cr, device=DEV) + 0.7
    gamma = gin(torch.rand(Cr, device=DEV) + 0.5)
    Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    ys, G, Gh = rnd(F_, Ho, Wo, C, seed=10), rnd(F_, H, W, C, seed=11), rnd(F_, Ho, Wo, C, seed=12)
    out, idx, ymax = run3(lambda: ops.pool_add_fwd(y, st.scale, st.shift, ys, st.scale, st.shift, want_ymax=True))
    idx, ymax = gin(idx), gin(ymax)
    run3(lambda: ops.bn_add_fwd(y, st.scale, st.shift, G))
    run3(lambda: ops.bn_add_fwd(y, st.scale, st.shift, G, st.scale, st.shift))
    run3(lambda: ops.bn_act(y, st.scale, st.shift, True))
    run3(lambda: ops.gather_s2(y))
    run3(lambda: ops.bn_relu_gap(y, st.scale, st.shift), exact=False)
    run3(lambda: ops.bn_bwd_sums(ymax, Gh), exact=False)
    dfeat = gin(torch.randn(F_, C, device=DEV))
    for training in (True, False):
        st.training = training

        def mk(mode, **kw):
            def f():
                dg = ops.torch.zeros((Cr,), device=DEV, dtype=F32); db = ops.torch.zeros((Cr,), device=DEV, dtype=F32)
                return ops.bn_bwd(mode, y, st, gamma, dg, db, **kw), dg, db
            return f
        # dy depends on the (atomically reduced) sums: summation-order noise times the cancellation of A*dz + B*y + C
        run3(mk(ops.SRC_DIRECT, G=G), exact=False, tol=1e-2)
        run3(mk(ops.SRC_RELU, G=G), exact=False, tol=1e-2)
        run3(mk(ops.SRC_POOL, G=Gh, idx=idx), exact=False, tol=1e-2)
        run3(mk(ops.SRC_GAP_RELU, dfeat=dfeat), exact=False, tol=1e-2)
        run3(mk(ops.SRC_RELU, G=G, grid_hw=(H + 2, W + 2)), exact=False, tol=1e-2)


@pytest.mark.parametrize("B,T,H", [(3, 7, 32), (4, 16, 128), (5, 6, 256), (2, 9, 512)])
def test_lstm_and_head(B, T, H):
    g = torch.Generator().manual_seed(13)
    xproj = gin((torch.randn(B * T, 4 * H, generator=g) * 0.5).to(DEV))
    b_ih, b_hh = gin(torch.randn(4 * H, generator=g).to(DEV) * 0.1), gin(torch.randn(4 * H, generator=g).to(DEV) * 0.1)
    w_hh = (torch.randn(4 * H, H, generator=g) * H ** -0.5).to(DEV)
    whh_b, whh_t = ops.pack_weight(w_hh, True)
    whh_b, whh_t = gin(whh_b), gin(whh_t)
    h, gates, cst, hn, cn = run3(lambda: ops.lstm_fwd(xproj, b_ih, b_hh, whh_t, B, T, H))
    h, gates, cst = gin(h), gin(gates), gin(cst)
    dout = gin(torch.randn(B, T, H, generator=g).to(DEV))

    def bwd():
        dbi = ops.torch.zeros((4 * H,), device=DEV, dtype=F32); dbh = ops.torch.zeros((4 * H,), device=DEV, dtype=F32)
        return ops.lstm_bwd(dout, None, None, gates, cst, h, whh_b, dbi, dbh, B, T, H)
    run3(bwd)
    a = gin(torch.randn(B, H, generator=g).to(DEV))
    W = gin((torch.randn(1024, H, generator=g) * H ** -0.5).to(DEV)); bias = gin(torch.randn(1024, generator=g).to(DEV))
    mask = gin((torch.rand(B, 1024, generator=g) > 0.3).to(torch.uint8).to(DEV))
    o = run3(lambda: ops.linear_small_fwd(a, W, bias, 1, mask, 1.0 / 0.7))[0]
    o = gin(o)
    delta = gin(torch.randn(B, 1024, generator=g).to(DEV))

    def lbwd():
        dW = ops.torch.zeros((1024, H), device=DEV, dtype=F32); db = ops.torch.zeros((1024,), device=DEV, dtype=F32)
        return ops.linear_small_bwd(delta, o, 1.0 / 0.7, a, W, dW, db), dW, db
    run3(lbwd, exact=False)
    z = gin(torch.randn(B, generator=g).to(DEV)); yb = gin(torch.randint(0, 2, (B,), generator=g).float().to(DEV))
    run3(lambda: ops.bce_fwd_bwd(z, yb, 0.1))
    p = gin(torch.rand(B, 1, generator=g).to(DEV))
    run3(lambda: ops.bce_prob_fwd_bwd(p, yb.view(B, 1)))


def test_layout_packing_and_misc():
    g = torch.Generator().manual_seed(14)
    x = gin(torch.randn(3, 728, 5, 7, generator=g).to(DEV))
    nhwc = run3(lambda: ops.nchw_to_nhwc(x, pad=True))[0]
    nhwc = gin(nhwc)
    back = run3(lambda: ops.nhwc_to_nchw(nhwc, 728))[0]
    assert (back - x).abs().max().item() < 0.02
    w = gin(torch.randn(728, 256, generator=g).to(DEV))
    run3(lambda: ops.pack_weight(w, True, pad=True))
    wd = gin(torch.randn(728, 1, 3, 3, generator=g).to(DEV))
    run3(lambda: ops.pack_dw(wd, pad=True))
    au = gin(torch.randn(6, 3, 13, 1, generator=g).to(DEV))
    run3(lambda: ops.bilinear_up(au, 64))
    run3(lambda: ops.cast_bf16(x))
    e = gin(torch.randn(5, 128, generator=g).to(DEV)); aw = gin(torch.randn(2, 128, generator=g).to(DEV) * 0.1)
    lab = gin(torch.tensor([0, 1, 1, 0, 1], device=DEV))

    def arc():
        dw = ops.torch.zeros((2, 128), device=DEV, dtype=F32)
        return ops.arcface_loss(e, aw, lab, 30.0, 0.3, 1, torch.tensor([1.8, 0.2], device=DEV), 2.0, dw=dw) + (dw,)
    run3(arc, exact=False)
    v, a = gin(torch.randn(3, 6, 256, generator=g).to(DEV)), gin(torch.randn(3, 6, 256, generator=g).to(DEV))
    run3(lambda: ops.fusion_pool_reg(v, a, 0.2, 0.1), exact=False)
