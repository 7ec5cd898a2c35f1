"""FusedAdam (one multi-tensor launch) against torch.optim.Adam / AdamW + clip_grad_norm_ -- the optimizer-side
mechanics of the reference loops (train_visual.py:533,574-577; train_au_face.py:616-619,678-693)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _params(dev, seed=0):
    g = torch.Generator().manual_seed(seed)
    shapes = [(3,), (17, 5), (728, 728, 1, 1), (1,), (64, 1, 3, 3), (8193,), (2048, 128)]
    return [torch.randn(*s, generator=g).to(dev) for s in shapes]


@pytest.mark.parametrize("decoupled,max_norm", [(False, None), (False, 1.0), (True, 0.5), (True, None)])
def test_fused_adam_matches_torch(decoupled, max_norm):
    from multimodal_deepfake_detection_b200 import FusedAdam
    dev = torch.device("cuda", 0)
    ours = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    ref = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    # eps = 1e-4 keeps the update a smooth function of the gradient: with the default 1e-8 the first steps are
    # lr * g / (|g| + eps), which amplifies the last-bit difference between fmaf(wd, p, g) and wd * p + g wherever g ~ 0
    o = FusedAdam(ours, lr=1e-2, eps=1e-4, weight_decay=1e-2, decoupled=decoupled, max_norm=max_norm)
    r = (torch.optim.AdamW if decoupled else torch.optim.Adam)(ref, lr=1e-2, eps=1e-4, weight_decay=1e-2)
    g = torch.Generator().manual_seed(5)
    for step in range(5):
        for a, b in zip(ours, ref):
            if step == 2 and a.numel() == 3:          # a parameter without a gradient is skipped by both
                a.grad = None; b.grad = None
                continue
            gr = torch.randn(a.shape, generator=g).to(dev) * (3.0 if step % 2 else 0.1)
            a.grad = gr.clone(); b.grad = gr.clone()
        if max_norm is not None:
            n_ref = torch.nn.utils.clip_grad_norm_([p for p in ref if p.grad is not None], max_norm)
        o.step(); r.step()
        if max_norm is not None:
            assert torch.allclose(o.grad_norm(), n_ref, rtol=1e-5)
        for a, b in zip(ours, ref):
            assert torch.allclose(a, b, rtol=2e-5, atol=2e-6), (step, a.shape, (a - b).abs().max().item())
    # state_dict has torch.optim.Adam's layout and round-trips
    sd = o.state_dict()
    sr = r.state_dict()
    assert set(sd["state"][0].keys()) == set(sr["state"][0].keys())
    assert float(sd["state"][1]["step"]) == 5.0
    assert not any(k.startswith("_xcp") for k in sd["param_groups"][0])
    ours2 = [torch.nn.Parameter(t.detach().clone()) for t in ours]
    o2 = FusedAdam(ours2, lr=1e-2, eps=1e-4, weight_decay=1e-2, decoupled=decoupled, max_norm=max_norm)
    o2.load_state_dict(sd)
    for a, a2, b in zip(ours, ours2, ref):
        gr = torch.randn(a.shape, generator=g).to(dev)
        a.grad = gr.clone(); a2.grad = gr.clone(); b.grad = gr.clone()
    if max_norm is not None:
        torch.nn.utils.clip_grad_norm_(ref, max_norm)
    o.step(); o2.step(); r.step()
    for a, a2, b in zip(ours, ours2, ref):
        assert torch.allclose(a, a2, rtol=1e-6, atol=1e-7)
        assert torch.allclose(a, b, rtol=2e-5, atol=2e-6)


def test_fused_adam_graph_capture():
    """The step is capturable: the step counter and the pointer table live in device memory."""
    from multimodal_deepfake_detection_b200 import FusedAdam
    dev = torch.device("cuda", 0)
    ours = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    ref = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    grads = [torch.randn_like(p) for p in ours]
    for p, q, g in zip(ours, ref, grads):
        p.grad = g.clone(); q.grad = g.clone()
    o = FusedAdam(ours, lr=1e-2, weight_decay=1e-3)
    r = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-3)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        o.step()
    torch.cuda.current_stream().wait_stream(s)
    r.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o.step()          # the capture itself does not execute
    for _ in range(3):
        graph.replay(); r.step()
    torch.cuda.synchronize()
    # the device-side step counters only move when the graph runs: 1 eager + 3 replayed steps
    assert int(o.state[ours[0]]["step"]) == 4
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=5e-5, atol=5e-6), (a - b).abs().max().item()


def test_fused_adam_graph_replay_follows_scheduler_and_state_dict_is_a_copy():
    """(1) lr / weight_decay live in device memory and a captured step re-reads them, so an LR scheduler keeps
    working under CUDA-graph replay; (2) state_dict() must not detach the live state from the arenas the graph updates:
    two exports taken around replays show the step count and the moments advancing."""
    from multimodal_deepfake_detection_b200 import FusedAdam
    dev = torch.device("cuda", 0)
    ours = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    ref = [torch.nn.Parameter(t.clone()) for t in _params(dev)]
    grads = [torch.randn_like(p) for p in ours]
    for p, q, g in zip(ours, ref, grads):
        p.grad = g.clone(); q.grad = g.clone()
    o = FusedAdam(ours, lr=1e-2, weight_decay=1e-3)
    r = torch.optim.Adam(ref, lr=1e-2, weight_decay=1e-3)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        o.step()
    torch.cuda.current_stream().wait_stream(s)
    r.step()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        o.step()
    sd0 = o.state_dict()
    for lr in (1e-2, 3e-3, 1e-3):                       # what ReduceLROnPlateau / OneCycleLR do between steps
        for grp in o.param_groups:
            grp["lr"] = lr
        for grp in r.param_groups:
            grp["lr"] = lr
        o.refresh_hyper()
        graph.replay(); r.step()
    torch.cuda.synchronize()
    for a, b in zip(ours, ref):
        assert torch.allclose(a, b, rtol=5e-5, atol=5e-6), (a - b).abs().max().item()
    sd1 = o.state_dict()
    assert float(sd0["state"][0]["step"]) == 1.0 and float(sd1["state"][0]["step"]) == 4.0
    assert not torch.equal(sd0["state"][0]["exp_avg"], sd1["state"][0]["exp_avg"])
    # the live state still aliases the arenas (int32 device counter, views of the flat moment buffers)
    st = o.state[ours[0]]
    assert st["step"].dtype == torch.int32 and int(st["step"]) == 4
    assert torch.equal(st["exp_avg"], sd1["state"][0]["exp_avg"])
    assert all(not k.startswith("_xcp_") for g in sd1["param_groups"] for k in g)
