"""Throughput of the GPU MFCC front-end (csrc/mfcc.cu) vs the numpy oracle on the host, same waveforms.

    python tools/mfcc_bench.py [--batch 64] [--frames 168]

168 frames = the 120/24/24 train/eval/test split the reference cuts from every file (wavfake_audio_dataset.py:8,64-70)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from multimodal_deepfake_detection_b200.audio_frontend import MFCC  # noqa: E402
from oracle import mfcc_oracle as M  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--frames", type=int, default=168)
    ap.add_argument("--iters", type=int, default=50)
    a = ap.parse_args()
    L = (a.frames - 1) * 160 + 80
    rng = np.random.default_rng(0)
    wav = (rng.standard_normal((a.batch, L)) * 0.1).astype(np.float32)
    dev = torch.device("cuda", 0)
    front = MFCC().to(dev)
    host = torch.from_numpy(wav).pin_memory()
    d = host.to(dev)
    for _ in range(5):
        front(d)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(a.iters):
        out = front(d)
    e1.record(); torch.cuda.synchronize()
    ms_dev = e0.elapsed_time(e1) / a.iters
    torch.cuda.synchronize(); e0.record()
    for _ in range(a.iters):
        out = front(host.to(dev, non_blocking=True)).cpu()
    e1.record(); torch.cuda.synchronize()
    ms_e2e = e0.elapsed_time(e1) / a.iters
    n_cpu = min(a.batch, 16)
    t0 = time.perf_counter()
    ref = [M.mfcc(wav[i]) for i in range(n_cpu)]
    cpu_ms = (time.perf_counter() - t0) * 1e3 / n_cpu * a.batch
    err = max(float(np.abs(out[i].numpy() - ref[i]).max()) for i in range(n_cpu))
    # algorithmic traffic: waveform read once + MFCC written once; the log-mel scratch is written and read once
    bytes_alg = a.batch * (L * 4 + a.frames * 13 * 4 + 2 * a.frames * 128 * 4)
    flops = a.batch * a.frames * (201 * 199 * 4 + 128 * 201 * 2 + 13 * 128 * 2)      # folded real DFT + mel + DCT
    print(json.dumps({"batch": a.batch, "frames": a.frames, "samples": L, "gpu_ms": ms_dev, "gpu_e2e_ms": ms_e2e,
                      "waveform_seconds_per_s": a.batch * L / 16000.0 / (ms_dev * 1e-3), "cpu_oracle_ms": cpu_ms, "cpu_cores": 1,
                      "speedup_e2e": cpu_ms / ms_e2e, "max_abs_err": err, "GB_per_s": bytes_alg / ms_dev / 1e6,
                      "GFLOP_per_s": flops / ms_dev / 1e6}))


if __name__ == "__main__":
    main()
