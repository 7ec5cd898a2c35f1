"""Size-independent properties at BASELINE.json's full shapes (16-frame 3x299x299 clips), where the fp32 oracle is too slow
to be the checker: batch independence and determinism in eval mode, permutation equivariance of the train-mode
forward, exactly-zero pad channels (728 -> 768 pitch), linearity of the backward pass in the upstream gradient."""
import warnings

import pytest
import torch

from multimodal_deepfake_detection_b200 import Xception, XceptionLSTMV, executor as ex, ops

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


@pytest.fixture(scope="module")
def net():
    torch.manual_seed(1234)
    return Xception(num_classes=2).to(DEV)


@pytest.fixture(scope="module")
def clip_frames():
    g = torch.Generator().manual_seed(5)
    return torch.rand(32, 3, 299, 299, generator=g).to(DEV)          # two 16-frame clips


def test_eval_forward_is_deterministic_and_batch_independent(net, clip_frames):
    net.eval()
    with torch.no_grad():
        full = net.features(clip_frames)
        again = net.features(clip_frames)
        halves = torch.cat([net.features(clip_frames[:16]), net.features(clip_frames[16:])])
        odd = torch.cat([net.features(clip_frames[:5]), net.features(clip_frames[5:])])
    assert torch.equal(full, again)                         # no atomics / split-K on the forward path
    assert torch.equal(full, halves) and torch.equal(full, odd)       # a frame's features do not depend on its batch (tiling)
    assert full.shape == (32, 2048) and torch.isfinite(full).all()


def test_train_forward_is_permutation_equivariant(net, clip_frames):
    """Batch statistics are sums over the batch: permuting the frames permutes the features (up to summation order)."""
    net.train()
    perm = torch.randperm(32, generator=torch.Generator().manual_seed(6)).to(DEV)
    with torch.no_grad():
        a = net.features(clip_frames)
        b = net.features(clip_frames[perm].contiguous())
    assert rel(b, a[perm]) < 2e-2
    assert int(net.bn1.num_batches_tracked) >= 2


def test_pad_channels_stay_exactly_zero(net, clip_frames):
    """728-channel tensors are stored with a 768 pitch; the 40 pad channels must be exact zeros in activations and
    gradients, otherwise they would leak into BatchNorm statistics."""
    net.train()
    x = clip_frames[:8]
    feat, tape = ex.xception_forward(net, x, save=True)
    mids = [bt for bt in tape.blocks if bt.out.shape[-1] == ops.phys(728)]
    assert len(mids) >= 9
    for bt in mids:
        assert bt.out.shape[-1] == 768
        assert not bt.out[..., 728:].any()
        for u in bt.units:
            if u.y.shape[-1] == 768:
                assert not u.y[..., 728:].any() and not u.st.scale[728:].any() and not u.st.shift[728:].any()
            if u.d.shape[-1] == 768:
                assert not u.d[..., 728:].any()
    params = net._backbone_params()
    sink = ex.GradSink(params, x.device, scratch_floats=2 * sum(ops.phys(p.numel()) + 4 for p in params if p.dim() == 1))
    ex.xception_backward(net, tape, torch.randn_like(feat), sink)
    for p in params:
        g = sink.view(p)
        assert g.shape == p.shape and torch.isfinite(g).all()
    assert sink.view(net.block5.rep[1].pointwise.weight).abs().sum() > 0


def test_backward_is_linear_in_the_upstream_gradient(net, clip_frames):
    """dL/dtheta is linear in dL/dfeat: backward(2 g) == 2 backward(g).  Scaling by 2 commutes with every rounding, so the
    only difference is the order of the fp32 RED accumulations (split-K weight gradients, BN sums).  Run with BatchNorm on
    running statistics: with batch statistics the BN-backward cancellations amplify that order noise to ~2e-2 on the
    smallest tensors even between two identical runs."""
    net.eval()
    x = clip_frames[:16]
    g = torch.randn(16, 2048, generator=torch.Generator().manual_seed(7)).to(DEV) * 1e-2
    grads = []
    for scale in (1.0, 2.0):
        net.zero_grad(set_to_none=True)
        f = net.features(x)
        f.backward(g * scale)
        grads.append({k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None})
    worst = max(rel(grads[1][k], 2.0 * grads[0][k]) for k in grads[0])
    assert worst < 1e-3, worst


def test_full_size_clip_step_smoke():
    """One 16x3x299x299 training step through the public modules: finite loss, every trainable parameter gets a gradient."""
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMV(128).to(DEV).train()
    for p in m.feature_extractor.parameters():
        p.requires_grad = True
    g = torch.Generator().manual_seed(8)
    clips = torch.rand(2, 16, 3, 299, 299, generator=g).to(DEV)
    y = torch.tensor([[1.0], [0.0]], device=DEV)
    loss = torch.nn.functional.binary_cross_entropy(m(m.extract_features(clips, torch.device(DEV))), y)
    loss.backward()
    assert torch.isfinite(loss)
    missing = [k for k, p in m.named_parameters() if p.requires_grad and p.grad is None]
    assert not missing, missing[:5]
    assert all(torch.isfinite(p.grad).all() for p in m.parameters() if p.grad is not None)


def test_800_frames_cross_the_32_bit_element_boundary(net):
    """Maximum sizes: 800 frames of 3x299x299 (50 clips per GPU) make the 147x147x128 activations 2.2e9 elements / 4.4 GB each,
    past 2^31 in elements and bytes.  Every frame is the same image, so (eval-mode BN) every feature row must equal the row
    computed from a 3-frame batch bit for bit, and with a per-frame-identical upstream gradient every parameter gradient
    must be (800 / 16) x the 16-frame gradient up to the order of the fp32 RED accumulations."""
    free, _ = torch.cuda.mem_get_info()
    if free < 150e9:
        pytest.skip("needs ~110 GB of free HBM")
    net.eval()
    g = torch.Generator().manual_seed(21)
    frame = torch.rand(1, 3, 299, 299, generator=g).to(DEV)
    grow = (torch.randn(1, 2048, generator=g) * 1e-2).to(DEV)
    with torch.no_grad():
        small = net.features(frame.expand(3, -1, -1, -1).contiguous())
        big = net.features(frame.expand(800, -1, -1, -1).contiguous())
    assert big.shape == (800, 2048)
    assert torch.equal(big, small[:1].expand(800, -1))

    def grads(n):
        net.zero_grad(set_to_none=True)
        f = net.features(frame.expand(n, -1, -1, -1).contiguous())
        f.backward(grow.expand(n, -1).contiguous())
        out = {k: p.grad.clone() for k, p in net.named_parameters() if p.grad is not None}
        net.zero_grad(set_to_none=True)
        return out
    g16 = grads(16)
    torch.cuda.empty_cache()
    g800 = grads(800)
    torch.cuda.empty_cache()
    errs = {k: rel(g800[k], 50.0 * g16[k]) for k in g16}
    worst = max(errs, key=errs.get)
    assert errs[worst] < 5e-3, (worst, errs[worst])
