"""Train-mode features with x2 materialised (round 1) vs read on the fly by block 1 (round 2) vs the fp32 oracle."""
import os, sys
import torch
sys.path.insert(0, ".")
from oracle import xception_oracle as O
from multimodal_deepfake_detection_b200 import Xception

DEV = "cuda"
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
sd = {k: v.to(DEV) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}
for H in (139, 299):
    g = torch.Generator().manual_seed(4)
    x = torch.rand(16, 3, H, H, generator=g).to(DEV)
    res = {}
    for mode in ("1", "0"):
        os.environ["XCP_MATERIALISE_X2"] = mode
        for training in (True, False):
            net = Xception(num_classes=2).to(DEV).train(training)
            net.load_state_dict(sd)
            with torch.no_grad():
                res[(mode, training)] = net.features(x)
    with torch.no_grad():
        for training in (True, False):
            ref = O.xception_features(sd, x, training, {})
            print("H=%d training=%d: materialised %.3e | fused %.3e | fused vs materialised %.3e" % (
                H, training, rel(res[("1", training)], ref), rel(res[("0", training)], ref), rel(res[("0", training)], res[("1", training)])))
