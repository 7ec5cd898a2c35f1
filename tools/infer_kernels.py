"""Per-kernel breakdown of the inference pass (eval-mode XceptionLSTMV forward, BN folded, no_grad):
CUDA events around every C-ABI call, grouped by bench.py's kernel families.  usage: python tools/infer_kernels.py [clips]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from multimodal_deepfake_detection_b200 import XceptionLSTMV, _lib  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.manual_seed(1234)
model = XceptionLSTMV(128).to(dev).eval()
clips = torch.rand(B, 16, 3, 299, 299, device=dev)
with torch.no_grad():
    for _ in range(3):
        model(model.extract_features(clips, dev))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        model(model.extract_features(clips, dev))
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    _lib.timer_start()
    for _ in range(2):
        model(model.extract_features(clips, dev))
    log = _lib.timer_stop()
kern, tot = bench.kernel_rooflines(log, ms, bench._peaks(), 2)
print("inference pass: %.3f ms for %d frames (%.0f frames/s); sum of kernels %.3f ms" % (ms, B * 16, B * 16 / ms * 1e3, tot))
for k in kern:
    print("%-55s %5.1f calls %8.1f us  share %.3f  frac %s" % (k["kernel"][:55], k["calls_per_step"], k["us_per_step"], k["share_of_step"],
                                                            "%.2f" % k["frac"] if k.get("frac") else "-"))
for (name, sig), (t, n) in sorted(log.items(), key=lambda kv: -kv[1][0])[:25]:
    print("   %-28s %8.1f us x %d  %s" % (name, t * 1e3 / n, n // 2, [a for a in sig if isinstance(a, int) and not isinstance(a, bool)][:9]))
