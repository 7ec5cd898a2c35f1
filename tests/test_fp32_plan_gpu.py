"""fp32 validation plan (csrc/f32.cu + csrc/f32_bwd.cu behind the SAME executor as the bf16 plan) against the UNMODIFIED
reference's golden vectors and the fp32 oracle, forward AND backward.

north_star tolerance: fp32 logits within 1e-4 relative.  `rel` is the relative L2 error; `relmax` the max-abs error over the
largest reference magnitude (stricter for single outliers)."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import xception_oracle as O  # noqa: E402
from multimodal_deepfake_detection_b200 import Xception, _lib, fp32_plan  # noqa: E402
from multimodal_deepfake_detection_b200._lib import XcpError  # noqa: E402

DEV = "cuda"
TOL = 1e-4


def setup_module(module):
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def relmax(a, b):
    return ((a.double() - b.double()).abs().max() / (b.double().abs().max() + 1e-30)).item()


def _t(a):
    return torch.from_numpy(np.asarray(a)).to(DEV)


@pytest.fixture(scope="module")
def sd2():
    return {k: v.to(DEV) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}


def _net(sd, train=False):
    net = Xception(num_classes=2).to(DEV)
    net.load_state_dict(sd)
    net.set_precision("fp32")
    return net.train() if train else net.eval()


# ---------------------------------------------------------------------------------------------- kernels one by one
@pytest.mark.parametrize("nchw,stride,ci,co,hw", [(True, 2, 3, 32, 37), (False, 1, 32, 64, 18)])
def test_f32_conv3x3(nchw, stride, ci, co, hw):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, ci, hw, hw + 2, generator=g).to(DEV)
    w = torch.randn(co, ci, 3, 3, generator=g).to(DEV)
    ref = F.conv2d(x.double(), w.double(), None, stride).permute(0, 2, 3, 1)
    out = fp32_plan.conv3x3(x if nchw else x.permute(0, 2, 3, 1).contiguous(), w, stride, nchw)
    assert out.shape == ref.shape and relmax(out, ref) < 2e-6


@pytest.mark.parametrize("shape", [(2, 19, 19, 728), (3, 7, 5, 64), (1, 1, 1, 8)])
def test_f32_dw3x3(shape):
    g = torch.Generator().manual_seed(2)
    x = torch.randn(*shape, generator=g).to(DEV)
    w = torch.randn(shape[3], 1, 3, 3, generator=g).to(DEV)
    ref = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, 1, 1, 1, groups=shape[3]).permute(0, 2, 3, 1)
    assert relmax(fp32_plan.dw3x3(x, w), ref) < 2e-6


@pytest.mark.parametrize("M,N,K,bias", [(722, 728, 728, False), (65, 2, 2048, True), (1, 1, 1, True), (300, 130, 37, False)])
def test_f32_gemm(M, N, K, bias):
    g = torch.Generator().manual_seed(3)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = torch.randn(N, K, generator=g).to(DEV)
    b = torch.randn(N, generator=g).to(DEV) if bias else None
    ref = a.double() @ w.double().t() + (b.double() if bias else 0.0)
    assert relmax(fp32_plan.gemm(a, w, b), ref) < 5e-6


def test_f32_pool_add_gather_gap():
    g = torch.Generator().manual_seed(4)
    y = torch.randn(2, 9, 11, 24, generator=g).to(DEV)
    s = torch.randn(2, 5, 6, 24, generator=g).to(DEV)
    out = torch.empty_like(s)
    _lib.call("xcp_f32_pool_add", fp32_plan._p(y), fp32_plan._p(s), fp32_plan._p(out), 2, 9, 11, 24, 0, fp32_plan._s())
    ref = F.max_pool2d(y.permute(0, 3, 1, 2), 3, 2, 1).permute(0, 2, 3, 1) + s
    assert torch.equal(out, ref)
    gth = torch.empty_like(s)
    _lib.call("xcp_f32_gather", fp32_plan._p(y), fp32_plan._p(gth), 2, 9, 11, 24, 2, 0, fp32_plan._s())
    assert torch.equal(gth, y[:, ::2, ::2, :])
    gap = torch.empty(2, 24, device=DEV)
    _lib.call("xcp_f32_gap", fp32_plan._p(y), fp32_plan._p(gap), 2, 99, 24, 0, fp32_plan._s())
    assert relmax(gap, y.double().mean(dim=(1, 2))) < 1e-6


# ---------------------------------------------------------------------------------------------- the network
def test_fp32_eval_matches_reference_golden(golden, sd2):
    """The reference's own fp32 outputs (oracle/gen_golden.py ran the unmodified reference): features at 299x299 and
    logits at 75x75 within 1e-4."""
    net = _net(sd2)
    with torch.no_grad():
        feat299 = net.features(_t(golden["x299"]))
        feat75 = net.features(_t(golden["x75"]))
        logits75 = net(_t(golden["x75"]))
    print("fp32 eval vs golden: feat299 %.2e feat75 %.2e logits75 %.2e" % (rel(feat299, _t(golden["A_eval_feat_299"])),
          rel(feat75, _t(golden["A_eval_feat_75"])), rel(logits75, _t(golden["A_eval_logits_75"]))))
    assert rel(feat299, _t(golden["A_eval_feat_299"])) < TOL
    assert rel(feat75, _t(golden["A_eval_feat_75"])) < TOL
    assert rel(logits75, _t(golden["A_eval_logits_75"])) < TOL and relmax(logits75, _t(golden["A_eval_logits_75"])) < TOL


def test_fp32_eval_logits_match_oracle_at_299(sd2):
    g = torch.Generator().manual_seed(11)
    x = torch.rand(4, 3, 299, 299, generator=g).to(DEV)
    net = _net(sd2)
    with torch.no_grad():
        ours = net(x)
        ref = O.xception_logits(sd2, x, False)
    print("fp32 eval vs oracle logits299: rel %.2e relmax %.2e" % (rel(ours, ref), relmax(ours, ref)))
    assert rel(ours, ref) < TOL and relmax(ours, ref) < TOL, (rel(ours, ref), relmax(ours, ref))


def test_fp32_train_mode_batch_statistics(golden, sd2):
    """Train-mode BatchNorm: batch statistics, running-stat updates and num_batches_tracked vs the reference's golden
    values (75x75, batch 2) and vs the oracle at 299x299."""
    net = _net(sd2, train=True)
    with torch.no_grad():
        logits = net(_t(golden["x75"]))
    print("fp32 train vs golden logits75: %.2e" % rel(logits, _t(golden["A_train_logits_75"])))
    assert rel(logits, _t(golden["A_train_logits_75"])) < TOL           # measured 1.3e-5 (batch of 2 at 3x3 resolution: BN amplifies rounding)
    cur = net.state_dict()
    for k in ("bn1", "block1.skipbn", "block4.rep.2", "block12.rep.5", "bn4"):
        assert rel(cur[k + ".running_mean"], _t(golden["A_rm::" + k])) < TOL, k
        assert rel(cur[k + ".running_var"], _t(golden["A_rv::" + k])) < TOL, k
        assert int(cur[k + ".num_batches_tracked"]) == int(golden["A_nbt::" + k])

    g = torch.Generator().manual_seed(12)
    x = torch.rand(6, 3, 299, 299, generator=g).to(DEV)
    net = _net(sd2, train=True)
    ns = {}
    with torch.no_grad():
        ours = net.features(x)
        ref = O.xception_features(sd2, x, True, ns)
    print("fp32 train vs oracle feat299: %.2e" % rel(ours, ref))
    assert rel(ours, ref) < TOL, rel(ours, ref)
    cur = net.state_dict()
    for k in ("bn1", "bn2", "block1.skipbn", "block3.rep.5", "block8.rep.8", "block12.rep.5", "bn3", "bn4"):
        assert rel(cur[k + ".running_mean"], ns[k + ".running_mean"]) < TOL, k
        assert rel(cur[k + ".running_var"], ns[k + ".running_var"]) < TOL, k


def test_fp32_uint8_frames_and_precision_switch(sd2):
    g = torch.Generator().manual_seed(13)
    u8 = torch.randint(0, 256, (2, 96, 96, 3), generator=g, dtype=torch.uint8).to(DEV)
    net = _net(sd2)
    with torch.no_grad():
        a = net.features(u8)
        b = O.xception_features(sd2, u8.permute(0, 3, 1, 2).float() / 255.0, False)
    assert rel(a, b) < TOL
    with pytest.raises(XcpError):
        net.set_precision("fp16")
    net.set_precision("bf16")
    with torch.no_grad():
        c = net.features(u8)
    assert 1e-7 < rel(c, b) < 2e-2          # back on the tensor-core plan


# ---------------------------------------------------------------------------------------------- Xception + LSTM + head
def _lstm_model(hidden, cls):
    import warnings
    feat_sd = O.synth_state_dict(1234, num_classes=None, bn_jitter=0.1)
    full = {"feature_extractor." + k: v for k, v in feat_sd.items()}
    full.update(O.synth_lstm_head_state_dict(77, hidden))
    full = {k: v.to(DEV) for k, v in full.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = cls(hidden).to(DEV)
    m.load_state_dict(full)
    return m.set_precision("fp32").eval(), full


def test_fp32_xception_lstmv_matches_reference_golden(golden):
    from multimodal_deepfake_detection_b200 import XceptionLSTMV
    m, _ = _lstm_model(32, XceptionLSTMV)
    with torch.no_grad():
        feats = m.extract_features(_t(golden["clips"]), torch.device(DEV))
        out = m.lstm(feats)[0]
        probs = m(feats)
    print("fp32 LSTMV vs golden: feats %.2e lstm_out %.2e probs %.2e" % (rel(feats, _t(golden["B_eval_feats"])),
          rel(out, _t(golden["B_eval_lstm_out"])), rel(probs, _t(golden["B_eval_probs"]))))
    assert rel(feats, _t(golden["B_eval_feats"])) < TOL
    assert rel(out, _t(golden["B_eval_lstm_out"])) < TOL
    assert rel(probs, _t(golden["B_eval_probs"])) < TOL
    with pytest.raises(XcpError):
        m.lstm(feats)                       # trainable LSTM parameters + grad enabled: forward-only plan refuses


def test_fp32_xception_lstma_matches_reference_golden(golden):
    from multimodal_deepfake_detection_b200 import XceptionLSTMA
    m, _ = _lstm_model(32, XceptionLSTMA)
    with torch.no_grad():
        feats = m.extract_features(_t(golden["C_audio"]), torch.device(DEV))
        probs = m(feats)
    print("fp32 LSTMA vs golden: feats %.2e probs %.2e" % (rel(feats, _t(golden["C_eval_feats"])), rel(probs, _t(golden["C_eval_probs"]))))
    assert rel(feats, _t(golden["C_eval_feats"])) < TOL
    assert rel(probs, _t(golden["C_eval_probs"])) < TOL


@pytest.mark.parametrize("H", [512, 1024])
def test_fp32_lstm_hidden_sizes_vs_torch(H):
    """The reference's hidden sizes (train_visual.py / train_audio.py): fp32 recurrence vs torch.nn.LSTM in float64."""
    from multimodal_deepfake_detection_b200.modules import FusedLSTM
    torch.manual_seed(5)
    ours = FusedLSTM(2048, H, 1, batch_first=True).to(DEV).set_precision("fp32")
    ref = torch.nn.LSTM(2048, H, 1, batch_first=True).to(DEV).double()
    ref.load_state_dict({k: v.double() for k, v in ours.state_dict().items()})
    x = torch.randn(3, 7, 2048, device=DEV)
    with torch.no_grad():
        o, (hn, cn) = ours(x)
        ro, (rhn, rcn) = ref(x.double())
    assert rel(o, ro) < 1e-5 and rel(hn, rhn) < 1e-5 and rel(cn, rcn) < 1e-5


# ---------------------------------------------------------------------------------------------- backward
def _leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
            for k, v in sd.items()}


@pytest.mark.parametrize("cfg", [(64, 128, 2, 2, False, True, 37), (128, 256, 2, 2, True, True, 22), (728, 728, 3, 1, True, True, 19),
                                  (728, 1024, 2, 2, True, False, 19), (64, 64, 2, 1, True, True, 12), (32, 48, 1, 1, True, True, 9)])
@pytest.mark.parametrize("training", [False, True])
def test_fp32_block_backward_every_tensor_within_1e4(cfg, training):
    """Block (all flavours of Xception.py:126-140 + a stride-1 skip-conv block) through the plan executor's backward on the fp32
    kernels: input gradient and EVERY parameter gradient within 1e-4 of the oracle's autograd, frozen and batch statistics."""
    from multimodal_deepfake_detection_b200 import Block
    cin, cout, reps, stride, swr, gf, hw = cfg
    torch.manual_seed(5)
    blk = Block(cin, cout, reps, stride, start_with_relu=swr, grow_first=gf).to(DEV).train(training).set_precision("fp32")
    with torch.no_grad():
        for m_ in blk.modules():
            if isinstance(m_, torch.nn.BatchNorm2d):
                m_.running_mean.normal_(0, 0.1); m_.running_var.uniform_(0.5, 1.5); m_.weight.uniform_(0.5, 1.5); m_.bias.normal_(0, 0.1)
    x = torch.randn(5, cin, hw, hw, device=DEV) * 0.7
    sd = {"b." + k: v.clone() for k, v in blk.state_dict().items()}
    leaves = _leaf(sd)
    xr = x.clone().requires_grad_(True)
    o = O.block_forward(leaves, "b", ("b", cin, cout, reps, stride, swr, gf), xr, training, {})
    dout = torch.randn_like(o)
    o.backward(dout)
    xo = x.clone().requires_grad_(True)
    out = blk(xo)
    out.backward(dout)
    errs = {"out": rel(out, o), "dx": rel(xo.grad, xr.grad)}
    errs.update({k: rel(p.grad, leaves["b." + k].grad) for k, p in blk.named_parameters()})
    worst = max(errs, key=errs.get)
    print("fp32 block %s training=%s: worst %s %.2e" % (cfg, training, worst, errs[worst]))
    assert errs[worst] < TOL, (worst, errs[worst])


def _xception_grads(sd2, training, n=6, seed=21, dtype64=False):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, 3, 299, 299, generator=g).to(DEV)
    labels = torch.randint(0, 2, (n,), generator=g).to(DEV)
    scale = 1.0 if training else 50.0
    net = _net(sd2, train=training)
    so = _leaf(sd2)
    lo = F.cross_entropy(O.xception_logits(so, x, training, {}) * scale, labels); lo.backward()
    l = F.cross_entropy(net(x) * scale, labels); l.backward()
    s64 = None
    if dtype64:
        s64 = {k: (v.double().clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else
                   (v.double().clone() if v.dtype.is_floating_point else v.clone())) for k, v in sd2.items()}
        F.cross_entropy(O.xception_logits(s64, x.double(), training, {}) * scale, labels).backward()
    return net, so, s64, l.item(), lo.item()


def test_fp32_xception_backward_every_tensor_within_1e4(sd2):
    """Whole backbone + fc, CE loss, 6x3x299x299 frames, frozen BatchNorm statistics: all 156 parameter gradients of the plan's
    backward within 1e-4 of the fp32 oracle (north_star asks 1e-2 of the bf16 plan; this pins the chain rule itself)."""
    net, so, _, l, lo = _xception_grads(sd2, False)
    errs = {k: rel(p.grad, so[k].grad) for k, p in net.named_parameters()}
    worst = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print("fp32 xception backward (frozen BN): loss %.6f vs %.6f, worst %s, median %.2e" % (
        l, lo, [(k, "%.2e" % v) for k, v in worst], float(np.median(list(errs.values())))))
    assert abs(l - lo) < 1e-5 * max(1.0, abs(lo))
    assert len(errs) == 156 and all(p.grad is not None for p in net.parameters())
    assert worst[0][1] < TOL, worst


def test_fp32_xception_backward_batch_statistics_vs_fp64_truth(sd2):
    """Train-mode BatchNorm (the mode the reference trains in, train_visual.py:558).  Through 40 stacked batch-statistics BNs the
    gradient of the seeded random-init network is ill-conditioned in fp32 ITSELF: the fp32 oracle differs from the fp64
    oracle by 4e-3 ... 8e-3 per tensor (tools/fp32_grad_probe.py; every Block alone, batch statistics included, is at 1e-6
    above).  So the truth here is the oracle in fp64, and the plan must be as close to it as the fp32 oracle is: per tensor
    within 2x the oracle's own fp32 error (+1e-4), same median, and the tensors above the first BN backward (fc, bn4.weight)
    within 1e-4 absolutely."""
    net, s32, s64, l, lo = _xception_grads(sd2, True, dtype64=True)
    ours = {k: rel(p.grad, s64[k].grad) for k, p in net.named_parameters()}
    o32 = {k: rel(s32[k].grad, s64[k].grad) for k in ours}
    worst = max(ours, key=lambda k: ours[k] / (o32[k] + 1e-4))
    print("fp32 xception backward (batch statistics) vs fp64 oracle: ours median %.2e max %.2e | fp32 oracle median %.2e max %.2e | "
          "worst ratio %s %.2e vs %.2e" % (float(np.median(list(ours.values()))), max(ours.values()),
                                          float(np.median(list(o32.values()))), max(o32.values()), worst, ours[worst], o32[worst]))
    assert abs(l - lo) < 1e-5
    for k in ("fc.weight", "fc.bias", "bn4.weight"):
        assert ours[k] < TOL, (k, ours[k])
    assert all(ours[k] < 2.0 * o32[k] + 1e-4 for k in ours), (worst, ours[worst], o32[worst])
    assert np.median(list(ours.values())) < 1.25 * np.median(list(o32.values())) + 1e-5


@pytest.mark.parametrize("M,K,N", [(361 * 3 + 5, 768, 768), (5476, 128, 256), (1000, 1536, 2048), (21609, 64, 128)])
def test_split3_makes_the_tensor_core_gemms_fp32_grade(M, K, N):
    """The PRODUCTION tcgen05 kernels (xcp_gemm_tn epi=2, xcp_gemm_wgrad) fed 3-way bf16 splits of fp32 operands reproduce the
    fp64 product to ~1e-6: indexing / accumulation / split-K logic of the kernels that train, checked far below bf16 noise."""
    from multimodal_deepfake_detection_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    b = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    out, _ = ops.gemm_tn(ops.split3(a, 0, False), ops.split3(b, 1, False), ops.EPI_F32)
    ref = (a.double() @ b.double().t())
    e_fwd = relmax(out, ref)
    plain, _ = ops.gemm_tn(a.to(torch.bfloat16), b.to(torch.bfloat16), ops.EPI_F32)
    dy = torch.randn(M, N, generator=g).to(DEV)
    dw = torch.zeros(N, K, device=DEV)
    ops.gemm_wgrad(ops.split3(dy, 0, True), ops.split3(a, 1, True), dw)
    ref_w = dy.double().t() @ a.double()
    e_w = relmax(dw, ref_w)
    print("split3 M=%d K=%d N=%d: gemm_tn %.2e (plain bf16 %.2e), wgrad %.2e" % (M, K, N, e_fwd, relmax(plain, ref), e_w))
    assert e_fwd < 2e-5 and e_w < 2e-5, (e_fwd, e_w)


def test_fp32_plan_on_split3_tensor_core_gemms(sd2):
    """Same backward as above with every pointwise / skip GEMM (forward, dgrad, wgrad) routed through the production tcgen05
    kernels on split operands (XCP_FP32_GEMM=split3)."""
    from multimodal_deepfake_detection_b200 import ops
    g = torch.Generator().manual_seed(22)
    x = torch.rand(3, 3, 299, 299, generator=g).to(DEV)
    labels = torch.randint(0, 2, (3,), generator=g).to(DEV)
    net = _net(sd2, train=False)
    so = _leaf(sd2)
    lo = F.cross_entropy(O.xception_logits(so, x, False, {}) * 50.0, labels); lo.backward()
    old = ops.FP32_GEMM[0]
    ops.FP32_GEMM[0] = "split3"
    try:
        l = F.cross_entropy(net(x) * 50.0, labels); l.backward()
    finally:
        ops.FP32_GEMM[0] = old
    errs = {k: rel(p.grad, so[k].grad) for k, p in net.named_parameters()}
    worst = max(errs, key=errs.get)
    print("fp32 plan on split3 tcgen05 GEMMs: worst %s %.2e median %.2e" % (worst, errs[worst], float(np.median(list(errs.values())))))
    assert errs[worst] < 5e-4, (worst, errs[worst])
