#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 hot path (contract in the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--clips B] [--impl ours|reference]

Workload (BASELINE.json `metric`): XceptionLSTMV(hidden_dim=128) *training* steps on synthetic clips of
16 x 3 x 299 x 299, backbone unfrozen (train_visual.py epochs >= 3: fwd + bwd through all 74 convs + LSTM + head
+ BCE + Adam).  One step = one pass over B clips per GPU.  `value` = clips/s over all GPUs with the inputs already
resident in HBM; `e2e` = the same step fed from pinned HOST buffers (H2D inside the timed region) with the loss
read back to the host every step.  N > 1: launched under torchrun, one rank per GPU, gradients averaged with
bucketed NCCL all-reduces overlapped with backward (weak scaling: B clips per GPU).

`--impl reference` times the reference's own algorithm on the host CPU cores (the fp32 oracle port in oracle/,
the reference itself cannot travel to the GPU box) on a bounded sample: one clip per step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "train clips/sec XceptionLSTMV 16x299x299"
T_FRAMES, HW, HIDDEN = 16, 299, 128


def _peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "_src": "fallback"}
    f = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(f):
        try:
            p.update(json.load(open(f)))
            p["_src"] = "measured"
        except Exception:
            pass
    return p


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
def cpu_oracle_clips_per_s(steps: int, warmup: int, threads: int):
    import torch
    import torch.nn.functional as F
    from oracle import xception_oracle as O

    torch.set_num_threads(threads)
    sd = O.synth_state_dict(1234, num_classes=None, bn_jitter=0.0)
    full = {"feature_extractor." + k: v for k, v in sd.items()}
    full.update(O.synth_lstm_head_state_dict(77, HIDDEN))
    leaves = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in full.items()}
    opt = torch.optim.Adam([v for v in leaves.values() if v.requires_grad], lr=1e-5, weight_decay=1e-4)
    g = torch.Generator().manual_seed(0)
    clips = torch.rand(1, T_FRAMES, 3, HW, HW, generator=g)
    y = torch.tensor([[1.0]])

    def step():
        opt.zero_grad(set_to_none=True)
        ns = {}
        prob = O.xception_lstm_forward(leaves, clips, training=True, new_stats=ns)
        loss = F.binary_cross_entropy(prob, y)
        loss.backward()
        opt.step()
        for k, v in ns.items():
            leaves[k] = v
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return steps / dt, dt / steps


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    cps, spc = cpu_oracle_clips_per_s(args.steps, args.warmup, threads)
    sample = "1 clip (16x3x299x299) per step: oracle fp32 fwd+bwd+Adam, backbone unfrozen, train-mode BN"
    line = {
        "impl": "reference", "metric": METRIC, "value": cps, "unit": "clips/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": spc * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "XceptionLSTMV(128) train step, backbone unfrozen, 16x3x299x299 clips", "clips_per_step": 1,
                   "device": "host CPU"},
        "cpu_baseline": {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": cps, "unit": "clips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# optional third arm (SURVEY.md §8d "the honest competitor"): the SAME module tree and weights driven through stock PyTorch
# (cuDNN / cuBLAS kernels, bf16 autocast, channels_last, fused torch Adam) on the same B200.  None of this package's kernels
# run here: the nn.Conv2d / nn.BatchNorm2d / nn.MaxPool2d / nn.LSTM / nn.Linear submodules execute their own torch forward.
def run_torch_stock(args):
    import warnings

    import torch
    import torch.nn as nn
    import torch.nn.functional as F

    from multimodal_deepfake_detection_b200 import SeparableConv2d, XceptionLSTMV

    if int(os.environ.get("RANK", "0")) != 0:
        return
    dev = torch.device("cuda", 0)
    torch.manual_seed(1234)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = XceptionLSTMV(HIDDEN).to(dev).train()
    for p in model.parameters():
        p.requires_grad = True
    model = model.to(memory_format=torch.channels_last)
    net = model.feature_extractor

    def sep(m, x):
        return m.pointwise(m.conv1(x))

    def block(b, inp):
        x = inp
        for m in b.rep:
            x = sep(m, x) if isinstance(m, SeparableConv2d) else m(x)
        return x + (b.skipbn(b.skip(inp)) if b.skip is not None else inp)

    def features(frames):
        x = F.relu(net.bn1(net.conv1(frames)))
        x = F.relu(net.bn2(net.conv2(x)))
        for i in range(1, 13):
            x = block(getattr(net, "block%d" % i), x)
        x = F.relu(net.bn3(sep(net.conv3, x)))
        x = F.relu(net.bn4(sep(net.conv4, x)))
        return F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)

    opt = torch.optim.Adam(model.parameters(), lr=1e-5, weight_decay=1e-4, fused=True)
    B = args.clips
    g = torch.Generator().manual_seed(0)
    clips = torch.rand(B, T_FRAMES, 3, HW, HW, generator=g).to(dev)
    y = torch.randint(0, 2, (B, 1), generator=g).float().to(dev)

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            frames = clips.view(B * T_FRAMES, 3, HW, HW).contiguous(memory_format=torch.channels_last)
            feats = features(frames).view(B, T_FRAMES, -1)
            out, _ = nn.LSTM.forward(model.lstm, feats)
        with torch.autocast("cuda", enabled=False):
            prob = model.sigmoid(model.fc_out(model.fc_layers(out[:, -1, :].float())))
            loss = F.binary_cross_entropy(prob, y)
        loss.backward()
        opt.step()
        return loss

    for _ in range(max(args.warmup, 3)):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    print(json.dumps({"impl": "torch-stock", "metric": METRIC, "value": B / (ms * 1e-3), "unit": "clips/s", "n_gpus": 1, "steps": args.steps,
                      "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "dtype": "bf16 autocast",
                      "data": "synthetic", "loss": float(loss.detach()),
                      "config": {"workload": "XceptionLSTMV(128) train step, backbone unfrozen, train-mode BN, BCE, torch.optim.Adam(fused)",
                                 "clips_per_gpu": B, "frames_per_clip": T_FRAMES, "frame": "3x299x299",
                                 "launch": "eager stock PyTorch %s, cuDNN %s, channels_last" % (torch.__version__, torch.backends.cudnn.version())}}),
          flush=True)


# ----------------------------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), "--query-gpu=" + self.QUERY,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            c = [x.strip() for x in r.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx = float(c[2])
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import torch.nn.functional as F

    from multimodal_deepfake_detection_b200 import FusedAdam, XceptionLSTMV, _lib, ops
    from multimodal_deepfake_detection_b200.ddp import GradBucketer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ops.check_device(dev)
    B = args.clips
    torch.manual_seed(1234)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        model = XceptionLSTMV(HIDDEN).to(dev)
    model.train()
    for p in model.feature_extractor.parameters():      # epoch >= freeze_epochs of train_visual.py:551-556
        p.requires_grad = True
    if world > 1:                                       # identical replicas
        for t in list(model.parameters()) + list(model.buffers()):
            dist.broadcast(t.data, 0)
    bucketer = GradBucketer(model, backbone=model.feature_extractor) if world > 1 else None
    opt = FusedAdam(model.parameters(), lr=1e-5, weight_decay=1e-4)      # Adam(lr=1e-5, weight_decay=1e-4), train_visual.py:533
    g = torch.Generator().manual_seed(1000 * rank)
    host_clips = [torch.rand(B, T_FRAMES, 3, HW, HW, generator=g).pin_memory() for _ in range(2)]
    host_y = [torch.randint(0, 2, (B, 1), generator=g).float().pin_memory() for _ in range(2)]
    dev_clips = host_clips[0].to(dev)
    dev_y = host_y[0].to(dev)

    from multimodal_deepfake_detection_b200 import BCELoss
    criterion = BCELoss()

    def step(clips, y):
        opt.zero_grad(set_to_none=True)
        feats = model.extract_features(clips, dev)
        prob = model(feats)
        loss = criterion(prob, y)                       # train_audio.py:20,39 criterion on the sigmoid output (one kernel)
        loss.backward()
        if bucketer is not None:
            bucketer.finish()
        opt.step()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            fn(i)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    W = max(args.warmup, 3)
    for i in range(W):
        step(dev_clips, dev_y)
    # ---- eager pass with CUDA events around every pointwise-GEMM launch (roofline) + launch census of one step
    torch.cuda.synchronize()
    _lib.reset_launch_count()
    ops.GEMM_TIMER.enable(True)
    n_eager = 2
    for i in range(n_eager):
        step(dev_clips, dev_y)
    launches_per_step = _lib.launch_count() // n_eager
    gemm_stats = ops.GEMM_TIMER.collect()
    ops.GEMM_TIMER.enable(False)

    # ---- the whole step (fwd, loss, bwd, all-reduce, Adam) captured once as a CUDA graph and replayed
    graphed = None
    mode = "eager"
    if not args.no_graph:
        try:
            from multimodal_deepfake_detection_b200.graph import GraphedTrainStep
            graphed = GraphedTrainStep(step, (dev_clips, dev_y), modules=[model], warmup=1)
            mode = "cuda-graph"
        except Exception as e:      # report and fall back to eager launches (still the same kernels)
            sys.stderr.write("bench.py: CUDA-graph capture failed (%s: %s); timing eager launches\n" % (type(e).__name__, e))
            graphed = None
            torch.cuda.synchronize()
    run = (lambda i: graphed.replay()) if graphed is not None else (lambda i: step(dev_clips, dev_y))
    for i in range(W):
        run(i)

    # ---- device-resident throughput
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms = timed(run, args.steps)
    launches = launches_per_step * args.steps
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / (ms * 1e-3)

    # ---- end to end: pinned host -> device every step (prefetched on a copy stream), loss read back every step
    from multimodal_deepfake_detection_b200.graph import HostPrefetcher
    pre = HostPrefetcher(dev)
    last = {}
    slot = {"k": pre.submit(host_clips[0], host_y[0])}

    def e2e_step(i):
        c, yy = pre.get(slot["k"])
        if graphed is not None:
            graphed.load_inputs(c, yy)
            slot["k"] = pre.submit(host_clips[(i + 1) & 1], host_y[(i + 1) & 1])     # next batch's H2D overlaps this step
            loss = graphed.replay()
        else:
            slot["k"] = pre.submit(host_clips[(i + 1) & 1], host_y[(i + 1) & 1])
            loss = step(c, yy)
        last["loss"] = float(loss.item())
    e2e_step(0)
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e * 1e-3)
    h2d = host_clips[0].numel() * 4 + host_y[0].numel() * 4

    def finish():
        # Leave without tearing the NCCL communicator down: destroy_process_group() can block on communicators that
        # were captured into the CUDA graph (seen at N=2: the JSON line printed, then the ranks never exited).
        sys.stdout.flush(); sys.stderr.flush()
        if world > 1:
            torch.cuda.synchronize()
            dist.barrier()
            torch.cuda.synchronize()
            os._exit(0)

    # ---- inference (the "infer frames/sec" half of the metric): eval-mode forward of the same clips, BN folded from the
    # running statistics, no saved activations (test_visual.py:609-624 protocol), device-resident inputs
    model.eval()
    def infer(i):
        with torch.no_grad():
            return model(model.extract_features(dev_clips, dev))
    for i in range(3):
        infer(i)
    ms_inf = timed(infer, args.steps)
    infer_fps = world * B * T_FRAMES * args.steps / (ms_inf * 1e-3)
    # latency of ONE clip (the per-video loop of test_visual.py:609-624): eager launches vs one CUDA-graph replay (row f-3)
    one_clip = None
    if rank == 0:
        from multimodal_deepfake_detection_b200.graph import GraphedInference
        clip1 = dev_clips[:1].contiguous()
        fwd1 = lambda c: model(model.extract_features(c, dev))  # noqa: E731

        def lat(fn, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record()
            for _ in range(n):
                fn()
            e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        with torch.no_grad():
            for _ in range(3):
                fwd1(clip1)
            ms_eager1 = lat(lambda: fwd1(clip1), 10)
        g1 = GraphedInference(fwd1, (clip1,), modules=[model])
        g1(clip1)
        one_clip = {"frames": T_FRAMES, "eager_ms": ms_eager1, "graph_ms": lat(lambda: g1(clip1), 20)}
        del g1
    model.train()

    # ---- the reference's FIRST training phase (train_visual.py:551-556, epochs < freeze_epochs): backbone frozen, BatchNorm
    # still in train mode (:558), only the LSTM + head get gradients.  Reported beside the headline (SURVEY.md §8d asks for both
    # modes); rank 0 only, eager launches, never allowed to disturb the line above.
    frozen = None
    if rank == 0:
        try:
            for p in model.feature_extractor.parameters():
                p.requires_grad = False
            head_params = [p for p in model.parameters() if p.requires_grad]
            opt_f = FusedAdam(head_params, lr=1e-5, weight_decay=1e-4)

            def frozen_step():
                opt_f.zero_grad(set_to_none=True)
                loss = criterion(model(model.extract_features(dev_clips, dev)), dev_y)
                loss.backward()
                opt_f.step()
            try:      # the graph capture above created AccumulateGrad nodes on its side stream; eager steps here only warn about it
                torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
            except AttributeError:
                pass
            for _ in range(3):
                frozen_step()
            ms_frozen = lat(frozen_step, max(args.steps, 3))
            frozen = {"value": B / (ms_frozen * 1e-3), "unit": "clips/s", "ms_per_step": ms_frozen, "n_gpus": 1,
                      "what": "same step with the backbone frozen (train-mode BN forward, LSTM + head trained), one GPU, eager"}
        except Exception as e:       # noqa: BLE001 - an auxiliary number must not take the bench line down
            frozen = {"error": "%s: %s" % (type(e).__name__, e)}
        finally:
            for p in model.feature_extractor.parameters():
                p.requires_grad = True

    if rank != 0:
        finish()
        return
    peaks = _peaks()
    # ---- roofline of the dominant kernel family: the middle-flow pointwise GEMM (M = F*361, K = N = 728) forward
    roof = None
    if gemm_stats:
        key = max(gemm_stats, key=lambda k: gemm_stats[k]["flops"])
        st = gemm_stats[key]
        ach = st["flops"] / (st["ms"] * 1e-3) / 1e12
        traffic = None                  # DRAM bytes per launch of this kernel from the committed ncu --set full capture
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r1_roofline_traffic.json")))
            if tj.get("shape") == key:
                traffic = tj["traffic_bytes_per_launch"]
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "gemm_kernel<256,EPI_BF16_STATS> (pointwise 1x1, %s)" % key, "achieved": ach,
                "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s", "frac": ach / peaks["bf16_tflops_sustained"],
                "traffic": traffic, "algorithmic_bytes": st["flops"] / st["n"] / (2.0 * 768) * 2 * 2 if "K=768 N=768" in key else None,
                "peak_source": peaks["_src"] + " (sustained: kernel timed inside a long step)",
                "launches_timed": st["n"], "avg_launch_us": st["ms"] * 1e3 / st["n"]}
    # ---- CPU baseline (bounded sample) on rank 0
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        cps, spc = cpu_oracle_clips_per_s(2, 1, threads)
        cpu = {"value": cps, "unit": "clips/s", "cores": threads, "kind": "port",
               "sample": "oracle fp32 port, 2 timed steps of 1 clip (16x3x299x299) fwd+bwd+Adam after 1 warm-up"}
    line = {
        "metric": METRIC, "value": value, "unit": "clips/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": "XceptionLSTMV(128) train step, backbone unfrozen, train-mode BN, BCE, fused Adam(1e-5, wd 1e-4)",
                   "clips_per_gpu": B, "global_batch": B * world, "frames_per_clip": T_FRAMES, "frame": "3x299x299",
                   "parallelism": "dp%d" % world, "launch": mode, "l2": "per-step working set (~%.0f GB of activations) >> 126 MB L2" % (B * 16 * 0.117)},
        "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": "clips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / max(args.steps, 1)},
        "gpu_launches": launches,
        "infer": {"value": infer_fps, "unit": "frames/s", "ms_per_pass": ms_inf / max(args.steps, 1),
                  "what": "XceptionLSTMV eval-mode forward (BN folded, no_grad), %d clips x %d frames per GPU per pass" % (B, T_FRAMES),
                  "one_clip_latency": one_clip},
        "frozen_backbone": frozen,
        "roofline": roof,
        "cpu_baseline": cpu,
        "loss": last.get("loss"),
    }
    print(json.dumps(line), flush=True)
    finish()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--clips", type=int, default=16, help="clips per GPU per step")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "torch-stock"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="time eager per-kernel launches instead of CUDA-graph replays")
    args = ap.parse_args()
    if args.impl == "torch-stock":
        run_torch_stock(args)
        return
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.gpus > 1 and world == 1:
        # convenience: re-launch under torchrun (the driver launches torchrun itself)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29517", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_ours(args)


if __name__ == "__main__":
    main()
