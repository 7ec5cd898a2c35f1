"""CUDA-graph replay of a whole training step must follow the same trajectory as eager launches of the same
kernels (graph.GraphedTrainStep), and the host prefetcher must deliver the submitted batches in order."""
import copy
import warnings

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from multimodal_deepfake_detection_b200 import XceptionLSTMV  # noqa: E402
from multimodal_deepfake_detection_b200.graph import GraphedTrainStep, HostPrefetcher  # noqa: E402

DEV = "cuda"


def _model(seed):
    torch.manual_seed(seed)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMV(32).to(DEV)
    m.train()
    for p in m.feature_extractor.parameters():
        p.requires_grad = True
    for mod in m.modules():                      # dropout masks come from different Philox offsets in eager / graph mode
        if isinstance(mod, torch.nn.Dropout):
            mod.eval()
    return m


def _make_step(m, opt):
    def step(clips, y):
        opt.zero_grad(set_to_none=True)
        loss = F.binary_cross_entropy(m(m.extract_features(clips, torch.device(DEV))), y)
        loss.backward()
        opt.step()
        return loss
    return step


def test_graphed_step_tracks_eager_steps():
    g = torch.Generator().manual_seed(5)
    batches = [(torch.rand(2, 3, 3, 75, 75, generator=g).to(DEV), torch.randint(0, 2, (2, 1), generator=g).float().to(DEV))
               for _ in range(4)]
    m_e = _model(7)
    m_g = copy.deepcopy(m_e)
    opt_e = torch.optim.Adam(m_e.parameters(), lr=1e-4, capturable=True)
    from multimodal_deepfake_detection_b200 import FusedAdam
    opt_g = FusedAdam(m_g.parameters(), lr=1e-4)          # one multi-tensor launch, device-side step counter: capturable
    step_e, step_g = _make_step(m_e, opt_e), _make_step(m_g, opt_g)
    # the capture warm-up executes real steps: give both models the same history
    warm = batches[0]
    for _ in range(2):                            # eager steps on the default stream BEFORE capture, as bench.py does:
        step_e(*warm)                             # their autograd graphs must not leak into the capture
        step_g(*warm)
    graphed = GraphedTrainStep(step_g, warm, modules=[m_g], warmup=1)
    step_e(*warm)
    losses_e, losses_g = [], []
    for clips, y in batches:
        losses_e.append(float(step_e(clips, y).detach()))
        losses_g.append(float(graphed(clips, y)))
    assert graphed.replays == len(batches)
    for a, b in zip(losses_e, losses_g):
        # bf16 kernels + RED-ordered weight gradients + 6-frame batch statistics: close, not bit-exact
        assert abs(a - b) < 2e-2, (losses_e, losses_g)
    # parameters moved, and moved alike
    w_e, w_g = m_e.fc_out.weight, m_g.fc_out.weight
    assert (w_e - w_g).abs().max().item() < 5e-3
    nbt = m_g.feature_extractor.bn1.num_batches_tracked.item()
    assert nbt == 3 + len(batches) and nbt == m_e.feature_extractor.bn1.num_batches_tracked.item()
    # eager evaluation after replays must see the replayed weights (pack caches are invalidated by replay())
    m_e.eval(); m_g.eval()
    with torch.no_grad():
        pe = m_e(m_e.extract_features(batches[0][0]))
        pg = m_g(m_g.extract_features(batches[0][0]))
    assert (pe - pg).abs().max().item() < 2e-2


def test_host_prefetcher_delivers_batches_in_order():
    pre = HostPrefetcher(torch.device(DEV))
    host = [torch.full((4, 1024), float(i)).pin_memory() for i in range(5)]
    k = pre.submit(host[0])
    for i in range(5):
        (t,) = pre.get(k)
        got = t.clone()
        if i + 1 < 5:
            k = pre.submit(host[i + 1])
        assert float(got.mean()) == float(i)


def test_graphed_inference_matches_eager_eval():
    """one clip's eval-mode forward replayed from a CUDA graph returns what eager launches return."""
    from multimodal_deepfake_detection_b200.graph import GraphedInference
    from multimodal_deepfake_detection_b200._lib import XcpError
    torch.manual_seed(3)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMV(32).to(DEV)
    g = torch.Generator().manual_seed(9)
    clips = [torch.rand(1, 4, 3, 96, 96, generator=g).to(DEV) for _ in range(3)]
    fn = lambda c: m(m.extract_features(c, torch.device(DEV)))  # noqa: E731
    with pytest.raises(XcpError):
        GraphedInference(fn, (clips[0],), modules=[m])          # still in train mode
    m.eval()
    with torch.no_grad():
        eager = [fn(c).clone() for c in clips]
    infer = GraphedInference(fn, (clips[0],), modules=[m])
    for c, e in zip(clips, eager):
        out = infer(c)
        assert torch.equal(out, e)
    assert infer.replays == 3
    # raw uint8 frames (row f-2) through the same capture path
    u8 = torch.randint(0, 256, (1, 4, 96, 96, 3), generator=g, dtype=torch.uint8).to(DEV)
    with torch.no_grad():
        e8 = fn(u8).clone()
    infer8 = GraphedInference(fn, (u8,), modules=[m])
    assert torch.equal(infer8(u8), e8)


def test_fusion_head_step_captures_after_eager_steps():
    """The audio-face fusion step (train_au_face.py:659-693 protocol: token streams -> FusionHead -> backward -> AdamW) captured
    after eager steps on the default stream, as bench.py --config c5 does.  Regression: an output tensor kept on the autograd ctx
    closed a reference cycle that kept the eager steps' graph (and the parameters' AccumulateGrad nodes on the default stream)
    alive, and the capture failed with cudaErrorStreamCaptureImplicit."""
    from multimodal_deepfake_detection_b200 import FusedAdam, FusionHead
    torch.manual_seed(3)
    head = FusionHead(64, samples_per_cls=(500, 10000), p_drop=0.0).to(DEV).train()
    proj_v = torch.nn.Linear(32, 64).to(DEV); proj_a = torch.nn.Linear(32, 64).to(DEV)      # stand-ins for the two token streams
    params = list(head.parameters()) + list(proj_v.parameters()) + list(proj_a.parameters())
    opt = FusedAdam(params, lr=1e-3, weight_decay=1e-2, decoupled=True, max_norm=1.0)
    g = torch.Generator().manual_seed(1)
    batches = [(torch.randn(4, 6, 32, generator=g).to(DEV), torch.randn(4, 9, 32, generator=g).to(DEV),
                torch.randint(0, 2, (4,), generator=g).to(DEV)) for _ in range(3)]

    def step(xv, xa, y):
        opt.zero_grad(set_to_none=True)
        loss, _ = head(proj_v(xv), proj_a(xa), y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(2):
        step(*batches[0])
    graphed = GraphedTrainStep(step, batches[0], modules=[head], warmup=1, optimizers=[opt])
    w_before = head.embed_head[0].weight.detach().clone()
    losses = [float(graphed(*b).detach()) for b in batches for _ in range(2)]
    assert all(torch.isfinite(torch.tensor(losses))) and graphed.replays == 6
    assert (head.embed_head[0].weight.detach() - w_before).abs().max().item() > 0        # the replayed steps trained
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
