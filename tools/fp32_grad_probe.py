"""Per-tensor gradient error of the fp32 validation plan against the oracle in fp32 AND fp64 (test infrastructure).

    python tools/fp32_grad_probe.py [frames] [--train] [--size N]

In train mode (batch statistics) the gradient of a random-init Xception is badly conditioned: this prints, per tensor, our
fp32 error and the fp32 oracle's own error against the fp64 oracle, in backward order (fc first)."""
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import Xception  # noqa: E402
from oracle import xception_oracle as O  # noqa: E402

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return ((a.double() - b.double()).norm() / (b.double().norm() + 1e-30)).item()


def leaf(sd, dt):
    return {k: (v.to(dt).clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else
                (v.to(dt).clone() if v.dtype.is_floating_point else v.clone())) for k, v in sd.items()}


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    n = int(args[0]) if args else 6
    train = "--train" in sys.argv
    size = int(sys.argv[sys.argv.index("--size") + 1]) if "--size" in sys.argv else 299
    sd = {k: v.to(DEV) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}
    g = torch.Generator().manual_seed(21)
    x = torch.rand(n, 3, size, size, generator=g).to(DEV); labels = torch.randint(0, 2, (n,), generator=g).to(DEV)
    scale = 1.0 if train else 50.0
    if "--contrast" in sys.argv:       # frames of the two classes differ strongly (brightness + texture scale): conditions train-mode BN
        lab = labels.view(n, 1, 1, 1).float()
        coarse = F.interpolate(torch.rand(n, 3, 30, 30, generator=g), size=(size, size), mode="bilinear").to(DEV)
        x = (lab * (0.55 + 0.4 * x) + (1 - lab) * (0.05 + 0.35 * coarse)).clamp(0, 1).contiguous()
    s32, s64 = leaf(sd, torch.float32), leaf(sd, torch.float64)
    l32 = F.cross_entropy(O.xception_logits(s32, x, train, {}) * scale, labels); l32.backward()
    l64 = F.cross_entropy(O.xception_logits(s64, x.double(), train, {}) * scale, labels); l64.backward()
    net = Xception(num_classes=2).to(DEV); net.load_state_dict(sd); net.set_precision("fp32"); net.train(train)
    l = F.cross_entropy(net(x) * scale, labels); l.backward()
    print("loss ours %.8f oracle32 %.8f oracle64 %.8f" % (l.item(), l32.item(), l64.item()))
    nb = Xception(num_classes=2).to(DEV); nb.load_state_dict(sd); nb.train(train)
    lb = F.cross_entropy(nb(x) * scale, labels); lb.backward()
    pb = dict(nb.named_parameters())
    eb = sorted(((rel(pb[k].grad, s64[k].grad), k) for k in pb), reverse=True)
    print("bf16 plan vs o64: loss %.6f worst %s median %.3e" % (lb.item(), [(k, "%.2e" % v) for v, k in eb[:3]], eb[len(eb) // 2][0]))
    names = [k for k, _ in net.named_parameters()][::-1]
    params = dict(net.named_parameters())
    print("%-36s %10s %10s %10s" % ("tensor", "ours-o64", "o32-o64", "ours-o32"))
    for k in names:
        print("%-36s %10.2e %10.2e %10.2e" % (k, rel(params[k].grad, s64[k].grad), rel(s32[k].grad, s64[k].grad), rel(params[k].grad, s32[k].grad)))


if __name__ == "__main__":
    main()
