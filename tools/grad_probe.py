"""Per-tensor gradient error of the bf16 production plan against the fp32 oracle (test infrastructure; imports oracle/).

    python tools/grad_probe.py [frames[:seed] ...] [--train] [--all] [--autocast]

Default protocol = tests/test_models_gpu.py::test_xception_gradients_*: Xception(num_classes=2), seeded weights, CE loss
x 50, frozen BN statistics (eval mode) unless --train.  --all prints every tensor, --autocast adds torch's own bf16-autocast
error on the same problem."""
import statistics
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, "."); sys.path.insert(0, "tests")
from multimodal_deepfake_detection_b200 import Xception  # noqa: E402
from oracle import xception_oracle as O  # noqa: E402

DEV = "cuda:0"
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-12)).item()


def leaf(sd):
    return {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}


def structured_frames(n, labels, g):
    """Low-frequency random fields plus a class-dependent pattern, in [0,1] (video_dataloader.py:35 range)."""
    low = F.interpolate(torch.randn(n, 3, 10, 10, generator=g), size=(299, 299), mode="bicubic", align_corners=False)
    mid = F.interpolate(torch.randn(n, 3, 40, 40, generator=g), size=(299, 299), mode="bilinear", align_corners=False)
    yy, xx = torch.meshgrid(torch.linspace(-1, 1, 299), torch.linspace(-1, 1, 299), indexing="ij")
    pat = torch.stack([torch.sin(6 * xx) * torch.cos(4 * yy), torch.cos(5 * xx + 3 * yy), torch.sin(7 * yy)])[None]
    s = (2.0 * labels.float() - 1.0).view(n, 1, 1, 1)
    return (0.5 + 0.18 * low + 0.08 * mid + 0.12 * s * pat + 0.02 * torch.randn(n, 3, 299, 299, generator=g)).clamp_(0, 1)


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    train, show_all, autocast = "--train" in sys.argv, "--all" in sys.argv, "--autocast" in sys.argv
    structured = "--structured" in sys.argv      # smooth, class-dependent frames instead of iid noise (conditions train-mode BN)
    cases = [(int(a.split(":")[0]), int(a.split(":")[1]) if ":" in a else 1) for a in args] or [(12, 1), (16, 1), (16, 3), (32, 5)]
    sd = {k: v.to(DEV) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}
    scale = 1.0 if train else 50.0
    for nfr, seed in cases:
        g = torch.Generator().manual_seed(seed)
        x = torch.rand(nfr, 3, 299, 299, generator=g).to(DEV); labels = torch.randint(0, 2, (nfr,), generator=g).to(DEV)
        if structured:
            x = structured_frames(nfr, labels.cpu(), g).to(DEV)
        net = Xception(num_classes=2).to(DEV); net.train(train); net.load_state_dict(sd); net.zero_grad(set_to_none=True)
        so = leaf(sd)
        lo = F.cross_entropy(O.xception_logits(so, x, train, {}) * scale, labels); lo.backward()
        l = F.cross_entropy(net(x) * scale, labels); l.backward()
        errs = {k: rel(p.grad, so[k].grad) for k, p in net.named_parameters()}
        w = sorted(errs.items(), key=lambda kv: -kv[1])
        print(nfr, seed, "train" if train else "frozen", "loss", round(l.item(), 4), round(lo.item(), 4), "worst",
              [(k, round(v, 4)) for k, v in w[:4]], "median", round(statistics.median(errs.values()), 4),
              "n>1e-2:", sum(1 for v in errs.values() if v > 1e-2), flush=True)
        if autocast:
            sb = leaf(sd)
            with torch.autocast("cuda", dtype=torch.bfloat16):
                lb = F.cross_entropy(O.xception_logits(sb, x, train, {}).float() * scale, labels)
            lb.backward()
            eb = {k: rel(sb[k].grad, so[k].grad) for k in errs}
            print("   torch autocast: worst", round(max(eb.values()), 4), "median", round(statistics.median(eb.values()), 4), flush=True)
        if show_all:
            for k, _ in net.named_parameters():
                print("   %-40s %.4e" % (k, errs[k]), flush=True)
        del net, so
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
