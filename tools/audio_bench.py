"""Config 4 of BASELINE.json (train_audio.py protocol): XceptionLSTMA(512) training step on (B, 120, 3, 13) MFCC clips,
backbone frozen (as constructed, XceptionLSTMA.py:10-12) or unfrozen, eager launches vs one CUDA-graph replay.

    python tools/audio_bench.py [--batch 8] [--steps 10] [--unfrozen]
"""
import argparse
import json
import os
import sys
import warnings

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_deepfake_detection_b200 import BCELoss, FusedAdam, XceptionLSTMA, _lib  # noqa: E402
from multimodal_deepfake_detection_b200.graph import GraphedTrainStep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--unfrozen", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="skip the CUDA-graph leg (for ncu launch lists)")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMA(512).to(dev).train()
    if a.unfrozen:
        for p in m.feature_extractor.parameters():
            p.requires_grad = True
    opt = FusedAdam([p for p in m.parameters()], lr=1e-4)
    crit = BCELoss()
    g = torch.Generator().manual_seed(1)
    mf = torch.randn(a.batch, a.frames, 1, 13, generator=g) * 20.0
    mf[..., 0] = mf[..., 0] * 5.0 - 300.0
    x = mf.repeat(1, 1, 3, 1).to(dev)
    y = torch.randint(0, 2, (a.batch, 1), generator=g).float().to(dev)

    def step(xx, yy):
        opt.zero_grad(set_to_none=True)
        loss = crit(m(m.extract_features(xx, dev)), yy)
        loss.backward()
        opt.step()
        return loss

    def timed(fn, n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(n):
            fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n
    for _ in range(3):
        step(x, y)
    _lib.reset_launch_count()
    step(x, y)
    launches = _lib.launch_count()
    ms_eager = timed(lambda: step(x, y), a.steps)
    ms_graph, last = ms_eager, step(x, y).detach()      # (a live loss would pin the eager autograd graph across the capture)
    if not a.no_graph:
        graphed = GraphedTrainStep(step, (x, y), modules=[m], warmup=1)
        last = graphed.replay()
        ms_graph = timed(graphed.replay, a.steps)
    m.eval()
    with torch.no_grad():
        for _ in range(2):
            m(m.extract_features(x, dev))
        ms_inf = timed(lambda: m(m.extract_features(x, dev)), a.steps)
    ms_inf_graph = None
    if not a.no_graph:
        from multimodal_deepfake_detection_b200.graph import GraphedInference
        gi = GraphedInference(lambda xx: m(m.extract_features(xx, dev)), (x,), modules=[m])
        gi(x)
        ms_inf_graph = timed(lambda: gi(x), a.steps)
    print(json.dumps({"workload": "XceptionLSTMA(512) train step, %s backbone" % ("unfrozen" if a.unfrozen else "frozen"),
                      "clips": a.batch, "frames_per_clip": a.frames, "launches_per_step": launches, "eager_ms": ms_eager,
                      "graph_ms": ms_graph, "clips_per_s_graph": a.batch / (ms_graph * 1e-3),
                      "patches_per_s_graph": a.batch * a.frames / (ms_graph * 1e-3), "infer_eager_ms": ms_inf, "infer_graph_ms": ms_inf_graph,
                      "infer_patches_per_s": a.batch * a.frames / (ms_inf * 1e-3), "loss": float(last.detach())}))


if __name__ == "__main__":
    main()
