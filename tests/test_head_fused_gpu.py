"""The classifier head + criterion as one launch per direction (csrc/head_fused.cu; BASELINE north_star (3)) against the
oracle's head (oracle/xception_oracle.py::head_forward, XceptionLSTMV.py:25-44,66-70) with injected dropout masks, the
reference's criteria (nn.BCELoss, train_audio.py:20,39; label-smoothing BCE-with-logits, train_au_patch.py:203-211), and
the unfused five-launch path it replaces."""
import os
import sys

import pytest
import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pytestmark = pytest.mark.gpu

from multimodal_deepfake_detection_b200 import ops  # noqa: E402
from multimodal_deepfake_detection_b200 import modules as M  # noqa: E402
from oracle import xception_oracle as orc  # noqa: E402

DEV = torch.device("cuda:0")


def rel_err(a, b):
    a, b = a.double().flatten(), b.double().flatten()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def head_params(H, Wd, seed):
    g = torch.Generator().manual_seed(seed)
    dims = [(Wd, H), (Wd, Wd), (Wd, Wd), (Wd, Wd), (1, Wd)]
    wb = []
    for n, k in dims:
        wb += [(torch.randn(n, k, generator=g) * (2.0 / k) ** 0.5).to(DEV), (torch.randn(n, generator=g) * 0.1).to(DEV)]
    return wb


def oracle_sd(wb):
    sd = {}
    for li, k in enumerate((0, 3, 6, 9)):
        sd["fc_layers.%d.weight" % k] = wb[2 * li]
        sd["fc_layers.%d.bias" % k] = wb[2 * li + 1]
    sd["fc_out.weight"], sd["fc_out.bias"] = wb[8], wb[9]
    return sd


@pytest.mark.parametrize("B,T,H,Wd,use_index", [(4, 16, 128, 1024, False), (16, 16, 128, 1024, True), (32, 5, 512, 1024, False),
                                                (1, 3, 128, 1024, False), (7, 4, 64, 256, True), (9, 2, 1024, 512, False)])
@pytest.mark.parametrize("mode", ["prob", "bce", "lsbce"])
def test_fused_head_matches_oracle_with_injected_masks(B, T, H, Wd, use_index, mode):
    g = torch.Generator().manual_seed(B * 131 + T)
    lstm_out = torch.randn(B, T, H, generator=g).to(DEV)
    wb = head_params(H, Wd, seed=H + Wd)
    keep = (torch.rand(4, B, Wd, generator=g) >= 0.3).to(DEV)
    y = torch.randint(0, 2, (B, 1), generator=g).float().to(DEV)
    row_index = torch.randint(0, T, (B,), generator=g).to(DEV) if use_index else None

    # ---- oracle: plain torch autograd on the reference's composition
    x_ref = lstm_out.clone().requires_grad_(True)
    wb_ref = [t.clone().requires_grad_(True) for t in wb]
    sel = x_ref[torch.arange(B, device=DEV), row_index] if use_index else x_ref[:, -1, :]
    h = sel
    for li in range(4):
        h = F.relu(F.linear(h, wb_ref[2 * li], wb_ref[2 * li + 1])) * keep[li].float() / 0.7
    z_ref = F.linear(h, wb_ref[8], wb_ref[9])
    p_ref = torch.sigmoid(z_ref)
    up = torch.randn(B, 1, generator=g).to(DEV)
    if mode == "prob":
        (p_ref * up).sum().backward()
    elif mode == "bce":
        loss_ref = torch.nn.BCELoss()(p_ref, y)
        (loss_ref * 1.7).backward()
    else:
        loss_ref = F.binary_cross_entropy_with_logits(z_ref, y * 0.9 + 0.05)
        (loss_ref * 1.7).backward()
    with torch.no_grad():      # and the oracle module's own head (same numbers, pins the composition to oracle/)
        p_orc = orc.head_forward(oracle_sd(wb), sel.detach(), [keep[li].float() / 0.7 for li in range(4)])
    assert rel_err(p_orc, p_ref.detach()) < 1e-6

    # ---- the fused kernels through the autograd functions the modules use
    x = lstm_out.clone().requires_grad_(True)
    wbp = [t.clone().requires_grad_(True) for t in wb]
    masks = keep.to(torch.uint8).contiguous()
    if mode == "prob":
        out = M._HeadFn.apply(x, row_index, True, 0.3, False, None, None, masks, *wbp)
        (out * up).sum().backward()
        assert rel_err(out, p_ref.detach()) < 1e-5
    else:
        out, loss = M._HeadLossFn.apply(x, row_index, True, 0.3, mode == "lsbce", y, 0.1 if mode == "lsbce" else None, masks, *wbp)
        (loss * 1.7).backward()
        assert abs(loss.item() - loss_ref.item()) < 1e-5 * max(1.0, abs(loss_ref.item()))
        assert rel_err(out, (z_ref if mode == "lsbce" else p_ref).detach()) < 1e-5
    assert rel_err(x.grad, x_ref.grad) < 1e-4
    for i, (a, b) in enumerate(zip(wbp, wb_ref)):
        assert rel_err(a.grad, b.grad) < 1e-4, (i, rel_err(a.grad, b.grad))
    bars, _ = ops.head_state(DEV)
    assert int(bars.abs().sum()) == 0           # both barrier counters back at rest


def test_fused_head_equals_the_five_launch_path_bitwise_forward():
    """Same per-neuron summation order as xcp_linear_small_fwd: the hidden activations are bit-identical."""
    B, T, H, Wd = 8, 16, 128, 1024
    g = torch.Generator().manual_seed(5)
    lstm_out = torch.randn(B, T, H, generator=g).to(DEV)
    wb = head_params(H, Wd, seed=9)
    masks = (torch.rand(4, B, Wd, generator=g) >= 0.3).to(torch.uint8).to(DEV)
    acts, z, prob, _, _ = ops.head_mlp_fwd(lstm_out, None, wb, 0.3, masks)
    a = lstm_out[:, -1, :].contiguous()
    for li in range(4):
        a = ops.linear_small_fwd(a, wb[2 * li], wb[2 * li + 1], 1, masks[li].contiguous(), 1.0 / 0.7)
        assert torch.equal(a, acts[li])
    assert rel_err(z, ops.linear_small_fwd(a, wb[8], wb[9], 0)) < 1e-6


def test_in_kernel_dropout_statistics_and_replay_counter():
    """Masks drawn in the kernel: keep rate 1-p per layer, fresh masks per launch (device-side counter, also under CUDA-graph
    replay), reproducible after a reseed, backward consistent with the drawn mask (finite-difference free check against
    autograd on the recovered mask)."""
    B, T, H, Wd = 16, 4, 128, 1024
    g = torch.Generator().manual_seed(11)
    lstm_out = torch.randn(B, T, H, generator=g).to(DEV)
    wb = head_params(H, Wd, seed=3)
    ops.head_reseed(DEV, 77)
    a1 = ops.head_mlp_fwd(lstm_out, None, wb, 0.3)[0]
    a2 = ops.head_mlp_fwd(lstm_out, None, wb, 0.3)[0]
    ops.head_reseed(DEV, 77)
    a3 = ops.head_mlp_fwd(lstm_out, None, wb, 0.3)[0]
    assert torch.equal(a1, a3) and not torch.equal(a1, a2)
    nodrop = ops.head_mlp_fwd(lstm_out, None, wb, 0.0)[0]
    alive = nodrop[0] > 0                                   # layer 0: same input with and without dropout
    kept = (a1[0] > 0)[alive].float().mean().item()
    assert abs(kept - 0.7) < 0.02, kept
    assert rel_err(a1[0][a1[0] > 0], (nodrop[0] / 0.7)[a1[0] > 0]) < 1e-6
    _, rng = ops.head_state(DEV)
    assert int(rng[1]) == 1                                 # one drawing launch since the reseed (p = 0 launches do not count)

    # CUDA-graph replays advance the device-side counter
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        ops.head_mlp_fwd(lstm_out, None, wb, 0.3)
    torch.cuda.current_stream().wait_stream(s)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ga = ops.head_mlp_fwd(lstm_out, None, wb, 0.3)[0]
    graph.replay(); r1 = ga.clone()
    graph.replay(); r2 = ga.clone()
    assert not torch.equal(r1, r2)

    # gradients against autograd with the mask recovered from the saved activations
    x = lstm_out.clone().requires_grad_(True)
    wbp = [t.clone().requires_grad_(True) for t in wb]
    y = torch.randint(0, 2, (B, 1), generator=g).float().to(DEV)
    ops.head_reseed(DEV, 5)
    out, loss = M._HeadLossFn.apply(x, None, True, 0.3, False, y, None, None, *wbp)
    loss.backward()
    ops.head_reseed(DEV, 5)
    acts = ops.head_mlp_fwd(lstm_out, None, wb, 0.3)[0]
    x_ref = lstm_out.clone().requires_grad_(True)
    wb_ref = [t.clone().requires_grad_(True) for t in wb]
    h = x_ref[:, -1, :]
    for li in range(4):
        pre = F.relu(F.linear(h, wb_ref[2 * li], wb_ref[2 * li + 1]))
        keep = ((acts[li] > 0) | (pre.detach() <= 0)).float()
        h = pre * keep / 0.7
    loss_ref = torch.nn.BCELoss()(torch.sigmoid(F.linear(h, wb_ref[8], wb_ref[9])), y)
    loss_ref.backward()
    assert abs(loss.item() - loss_ref.item()) < 1e-5
    assert rel_err(x.grad, x_ref.grad) < 1e-4
    for a, b in zip(wbp, wb_ref):
        assert rel_err(a.grad, b.grad) < 1e-4


def test_model_forward_loss_equals_forward_plus_criterion():
    """XceptionLSTMV.forward_loss == BCELoss()(forward(.), y) and XceptionLSTMA-style forward_logits + label-smoothing BCE,
    values and every gradient (dropout off so both paths see the same network)."""
    torch.manual_seed(1234)
    model = M.XceptionLSTMV(128).to(DEV).train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    g = torch.Generator().manual_seed(3)
    feats = torch.randn(6, 16, 2048, generator=g).to(DEV)
    y = torch.randint(0, 2, (6, 1), generator=g).float().to(DEV)
    params = [p for n, p in model.named_parameters() if not n.startswith("feature_extractor")]

    def grads(fn):
        for p in params:
            p.grad = None
        loss, out = fn()
        loss.backward()
        return loss.detach().clone(), out.detach().clone(), [p.grad.clone() for p in params]

    l1, o1, g1 = grads(lambda: model.forward_loss(feats, y))
    def two_call():
        out = model(feats)
        return M.BCELoss()(out, y), out
    l2, o2, g2 = grads(two_call)
    assert abs(l1.item() - l2.item()) < 1e-6 and rel_err(o1, o2) < 1e-6
    for a, b in zip(g1, g2):
        assert rel_err(a, b) < 1e-4
    l3, o3, g3 = grads(lambda: model.forward_loss(feats, y, smoothing=0.1))
    def two_call_logits():
        out = model.forward_logits(feats)
        return M.LabelSmoothingBCEWithLogitsLoss(0.1)(out, y), out
    l4, o4, g4 = grads(two_call_logits)
    assert abs(l3.item() - l4.item()) < 1e-6 and rel_err(o3, o4) < 1e-6
    for a, b in zip(g3, g4):
        assert rel_err(a, b) < 1e-4
    # against the reference's own criterion on the same probabilities
    assert abs(l1.item() - torch.nn.BCELoss()(o1, y).item()) < 1e-6


@pytest.mark.parametrize("B,Tv,Ta,D,focal", [(4, 16, 16, 256, True), (9, 16, 120, 256, True), (32, 5, 7, 128, False), (1, 1, 3, 64, True),
                                             (6, 2, 1, 512, False)])
def test_fused_fusion_head_matches_oracle(B, Tv, Ta, D, focal):
    """train_au_face.py:659-674 as one launch per direction against oracle/xception_oracle.py::fusion_head_loss under torch
    autograd: loss, logits and every gradient (both token streams of different lengths, embed_head, ArcFace weight), with an
    injected dropout mask and a non-unit upstream gradient."""
    g = torch.Generator().manual_seed(B * 17 + D)
    v = torch.randn(B, Tv, D, generator=g).to(DEV); a = torch.randn(B, Ta, D, generator=g).to(DEV)
    W0 = (torch.randn(256, 2 * D, generator=g) * (1.0 / D) ** 0.5).to(DEV); b0 = (torch.randn(256, generator=g) * 0.1).to(DEV)
    W3 = (torch.randn(128, 256, generator=g) * (2.0 / 256) ** 0.5).to(DEV); b3 = (torch.randn(128, generator=g) * 0.1).to(DEV)
    arc = torch.randn(2, 128, generator=g).to(DEV)
    labels = torch.randint(0, 2, (B,), generator=g).to(DEV)
    keep = (torch.rand(B, 256, generator=g) >= 0.2).to(DEV)
    cw = orc.cb_focal_weights([500, 10000]).to(DEV) if focal else None

    leaves = [t.clone().requires_grad_(True) for t in (v, a, W0, b0, W3, b3, arc)]
    rv, ra, rW0, rb0, rW3, rb3, rarc = leaves
    if focal:
        loss_ref, logits_ref = orc.fusion_head_loss({"0.weight": rW0, "0.bias": rb0, "3.weight": rW3, "3.bias": rb3}, rarc, rv, ra, labels, cw,
                                                    s=30.0, m=0.30, gamma=2.0, lambda_align=0.2, lambda_temp=0.1, drop_mask=keep.float() / 0.8)
    else:       # plain cross entropy on the margin logits (train_visual.py:455-474,532) + the same regularisers
        vp, ap = rv.mean(1), ra.mean(1)
        hdn = F.relu(F.linear(torch.cat([vp, ap], 1), rW0, rb0)) * keep.float() / 0.8
        logits_ref = orc.arcface_logits(rarc, F.linear(hdn, rW3, rb3), labels, 30.0, 0.30)
        ltv = (rv[:, 1:] - rv[:, :-1]).pow(2).mean() if Tv > 1 else 0.0
        lta = (ra[:, 1:] - ra[:, :-1]).pow(2).mean() if Ta > 1 else 0.0
        loss_ref = F.cross_entropy(logits_ref, labels) + 0.2 * F.mse_loss(vp, ap) + 0.1 * 0.5 * (ltv + lta)
    (loss_ref * 1.3).backward()

    mine = [t.clone().requires_grad_(True) for t in (v, a, W0, b0, W3, b3, arc)]
    cfg = (30.0, 0.30, 2.0, 0.2, 0.1, True, 0.2)
    loss, logits = M._FusionHeadFn.apply(mine[0], mine[1], labels, mine[2], mine[3], mine[4], mine[5], mine[6], cw, cfg,
                                         keep.to(torch.uint8).contiguous())
    (loss * 1.3).backward()
    assert abs(loss.item() - loss_ref.item()) < 2e-5 * max(1.0, abs(loss_ref.item()))
    assert rel_err(logits, logits_ref.detach()) < 1e-5
    names = ("v_tokens", "au_tokens", "embed.0.weight", "embed.0.bias", "embed.3.weight", "embed.3.bias", "arcface.weight")
    for n, x, r in zip(names, mine, leaves):
        assert rel_err(x.grad, r.grad) < 2e-4, (n, rel_err(x.grad, r.grad))
    bars, _ = ops.head_state(DEV)
    assert int(bars.abs().sum()) == 0


def test_fusion_head_module_inference_logits_and_dropout_draw():
    """FusionHead.predict_logits == s * cos of the embedding (labels=None path, train_au_face.py:715-716); train-mode dropout is drawn
    in the kernel (two calls differ, keep rate ~0.8)."""
    torch.manual_seed(5)
    head = M.FusionHead(256, samples_per_cls=(500, 10000)).to(DEV)
    g = torch.Generator().manual_seed(9)
    v = torch.randn(8, 16, 256, generator=g).to(DEV); a = torch.randn(8, 16, 256, generator=g).to(DEV)
    y = torch.randint(0, 2, (8,), generator=g).to(DEV)
    head.eval()
    lg = head.predict_logits(v, a)
    e = head.embed_head
    emb = F.linear(F.relu(F.linear(torch.cat([v.mean(1), a.mean(1)], 1), e[0].weight, e[0].bias)), e[3].weight, e[3].bias)
    ref = orc.arcface_logits(head.arcface.weight, emb, None, head.arcface.s, head.arcface.m)
    assert rel_err(lg, ref.detach()) < 1e-5
    head.train()
    l1, _ = head(v, a, y)
    l2, _ = head(v, a, y)
    assert torch.isfinite(l1) and torch.isfinite(l2) and l1.item() != l2.item()
    l1.backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in head.parameters())


def test_head_on_more_than_32_clips_runs_in_chunks():
    """More clips than one head launch holds (32): forward / forward_logits / forward_loss split the batch over several launches;
    values and gradients equal the oracle's head on the whole batch."""
    torch.manual_seed(11)
    model = M.XceptionLSTMV(128).to(DEV).train()
    for mod in model.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    g = torch.Generator().manual_seed(2)
    feats = torch.randn(45, 4, 2048, generator=g).to(DEV)
    y = torch.randint(0, 2, (45, 1), generator=g).float().to(DEV)
    params = [p for n, p in model.named_parameters() if n.startswith("fc_")]
    loss, probs = model.forward_loss(feats, y)
    for p in params:
        p.grad = None
    loss.backward()
    got = [p.grad.clone() for p in params]
    with torch.no_grad():
        lstm_out, _ = model.lstm(feats)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in model.state_dict().items() if k.startswith("fc_")}
    p_ref = orc.head_forward(sd, lstm_out[:, -1].float())
    loss_ref = torch.nn.BCELoss()(p_ref, y)
    loss_ref.backward()
    assert probs.shape == (45, 1) and rel_err(probs, p_ref.detach()) < 1e-5 and abs(loss.item() - loss_ref.item()) < 1e-5
    for (n, _), a in zip([(n, p) for n, p in model.named_parameters() if n.startswith("fc_")], got):
        assert rel_err(a, sd[n].grad) < 1e-4, n
    assert rel_err(model.forward_logits(feats), orc.head_forward(sd, lstm_out[:, -1].float(), return_logit=True).detach()) < 1e-5
