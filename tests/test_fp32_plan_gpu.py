"""fp32 validation plan (csrc/f32.cu + fp32_plan.py) against the UNMODIFIED reference's golden vectors and the fp32 oracle.

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


def test_fp32_uint8_frames_and_forward_only_guard(sd2):
    g = torch.Generator().manual_seed(13)
    u8 = torch.randint(0, 256, (2, 96, 96, 3), generator=g, dtype=torch.uint8).to(DEV)
    net = _net(sd2)
    with torch.no_grad():
        a = net.features(u8)
        b = O.xception_features(sd2, u8.permute(0, 3, 1, 2).float() / 255.0, False)
    assert rel(a, b) < TOL
    with pytest.raises(XcpError):
        net.features(u8)                    # grad enabled + trainable parameters: the fp32 plan refuses (forward-only)
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
