import sys, statistics, torch
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch.nn.functional as F
from multimodal_deepfake_detection_b200 import Xception
from oracle import xception_oracle as O
DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
sd = {k: v.to(DEV) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}
def rel(a, b): return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()
for nfr, seed in [(12, 1), (16, 1), (16, 3), (32, 5)]:
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(nfr, 3, 299, 299, generator=g).to(DEV); labels = torch.randint(0, 2, (nfr,), generator=g).to(DEV)
    net = Xception(num_classes=2).to(DEV); net.train(False); net.load_state_dict(sd); net.zero_grad(set_to_none=True)
    so = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    lo = F.cross_entropy(O.xception_logits(so, x, False, {}) * 50.0, labels); lo.backward()
    l = F.cross_entropy(net(x) * 50.0, labels); l.backward()
    errs = {k: rel(p.grad, so[k].grad) for k, p in net.named_parameters()}
    w = sorted(errs.items(), key=lambda kv: -kv[1])[:3]
    print(nfr, seed, "loss", round(l.item(), 4), round(lo.item(), 4), "worst", [(k, round(v, 4)) for k, v in w], "median", round(statistics.median(errs.values()), 4), flush=True)
