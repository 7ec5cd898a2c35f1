"""Diagnostic (GPU): our Xception / XceptionLSTMV against the fp32 oracle on identical seeded weights and inputs.
Prints per-tensor errors and, beside them, the error of the oracle itself under torch bf16 autocast (the
reference's own bf16 noise floor, SURVEY.md §7 hard-part 4).  Usage: python tools/model_parity.py [F] [H]"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, ".")
from oracle import xception_oracle as O  # noqa: E402
from multimodal_deepfake_detection_b200 import Xception, XceptionLSTMV  # noqa: E402

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = "cuda"
Fr = int(sys.argv[1]) if len(sys.argv) > 1 else 8
H = int(sys.argv[2]) if len(sys.argv) > 2 else 139


def rel(a, b):
    return ((a.float() - b.float()).norm() / (b.float().norm() + 1e-20)).item()


def relmax(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-20)).item()


sd = {k: v.to(dev) for k, v in O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1).items()}
g = torch.Generator().manual_seed(0)
x = torch.rand(Fr, 3, H, H, generator=g).to(dev)
labels = torch.randint(0, 2, (Fr,), generator=g).to(dev)

net = Xception(num_classes=2).to(dev)
net.load_state_dict(sd)

for mode in ("eval", "train"):
    training = mode == "train"
    net.train(training)
    # ---- oracle fp32
    sdo = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
    ns = {}
    feat_o = O.xception_features(sdo, x, training, ns)
    logit_o = F.linear(feat_o, sdo["fc.weight"], sdo["fc.bias"])
    # ---- oracle under bf16 autocast (noise floor of the reference itself)
    with torch.autocast("cuda", dtype=torch.bfloat16):
        sdb = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in sd.items()}
        feat_b = O.xception_features(sdb, x, training, {})
        logit_b = F.linear(feat_b, sdb["fc.weight"], sdb["fc.bias"])
    # ---- ours
    net.load_state_dict(sd)
    net.zero_grad(set_to_none=True)
    torch.cuda.synchronize(); t0 = time.time()
    fc = net.fc
    feat = net.features(x)
    logit = fc(feat)
    torch.cuda.synchronize(); t1 = time.time()
    print(f"[{mode}] F={Fr} H={H} ours fwd {1e3*(t1-t0):.1f} ms")
    print(f"[{mode}] feat   rel-L2 ours {rel(feat, feat_o):.3e}  max-rel {relmax(feat, feat_o):.3e} | autocast-bf16 oracle {rel(feat_b, feat_o):.3e} {relmax(feat_b, feat_o):.3e}")
    print(f"[{mode}] logits rel-L2 ours {rel(logit, logit_o):.3e} max-rel {relmax(logit, logit_o):.3e} | autocast-bf16 oracle {rel(logit_b, logit_o):.3e}")
    if not training:
        # gradient parity with frozen BN statistics (well conditioned: no batch-stat amplification)
        lo = F.cross_entropy(logit_o * 50, labels); lo.backward()
        lb = F.cross_entropy(logit_b.float() * 50, labels); lb.backward()
        l = F.cross_entropy(logit * 50, labels); l.backward()
        worst = sorted(((rel(p.grad, sdo[k].grad), rel(sdb[k].grad, sdo[k].grad), k) for k, p in net.named_parameters()), reverse=True)
        import statistics
        print("   [eval-BN] grad rel-L2 per tensor: median ours %.3e (bf16-oracle %.3e); max ours %.3e (%s)" % (
            statistics.median(w[0] for w in worst), statistics.median(w[1] for w in worst), worst[0][0], worst[0][2]))
        for e, eb, k in worst[:8]:
            print(f"     {k:40s} ours {e:.3e}   bf16-oracle {eb:.3e}")
    if training:
        for k in ("bn1", "block1.skipbn", "block4.rep.2", "block12.rep.5", "bn4"):
            print(f"   running_mean {k}: {rel(net.state_dict()[k + '.running_mean'], ns[k + '.running_mean']):.2e}  running_var: "
                  f"{rel(net.state_dict()[k + '.running_var'], ns[k + '.running_var']):.2e}  nbt {int(net.state_dict()[k + '.num_batches_tracked'])}")
        loss_o = F.cross_entropy(logit_o, labels); loss_o.backward()
        loss_b = F.cross_entropy(logit_b.float(), labels); loss_b.backward()
        loss = F.cross_entropy(logit, labels)
        torch.cuda.synchronize(); t0 = time.time()
        loss.backward()
        torch.cuda.synchronize(); t1 = time.time()
        print(f"[train] loss ours {loss.item():.6f} oracle {loss_o.item():.6f} bf16-oracle {loss_b.item():.6f}; ours bwd {1e3*(t1-t0):.1f} ms")
        worst = []
        for k, p in net.named_parameters():
            e = rel(p.grad, sdo[k].grad); eb = rel(sdb[k].grad, sdo[k].grad)
            worst.append((e, eb, k))
        worst.sort(reverse=True)
        import statistics
        print("   grad rel-L2 per tensor: median ours %.3e (bf16-oracle %.3e); max ours %.3e" % (
            statistics.median(w[0] for w in worst), statistics.median(w[1] for w in worst), worst[0][0]))
        for e, eb, k in worst[:12]:
            print(f"     {k:40s} ours {e:.3e}   bf16-oracle {eb:.3e}")
        for k in ("conv1.weight", "conv2.weight", "bn1.weight", "block1.rep.0.conv1.weight", "block1.skip.weight", "fc.weight"):
            e = rel(dict(net.named_parameters())[k].grad, sdo[k].grad)
            print(f"     {k:40s} ours {e:.3e}")

# ---- XceptionLSTMV end to end (BCE path, dropout off)
hid = 128
m = XceptionLSTMV(hid).to(dev)
feat_sd = {("feature_extractor." + k): v for k, v in sd.items() if not k.startswith("fc.")}
lh = {k: v.to(dev) for k, v in O.synth_lstm_head_state_dict(77, hid).items()}
full = dict(feat_sd); full.update(lh)
m.load_state_dict(full)
B, T = max(Fr // 4, 1), 4
clips = torch.rand(B, T, 3, H, H, generator=g).to(dev)
y = torch.randint(0, 2, (B, 1), generator=g).float().to(dev)
m.train()
for mod in m.modules():
    if isinstance(mod, torch.nn.Dropout):
        mod.eval()
for p in m.feature_extractor.parameters():
    p.requires_grad = True
fo = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in full.items()}
prob_o = O.xception_lstm_forward(fo, clips, training=True, new_stats={})
loss_o = F.binary_cross_entropy(prob_o, y); loss_o.backward()
feats = m.extract_features(clips, torch.device(dev))
prob = m(feats)
loss = F.binary_cross_entropy(prob, y)
loss.backward()
print(f"[lstmv] probs ours {prob.flatten().tolist()} oracle {prob_o.flatten().tolist()}")
print(f"[lstmv] loss ours {loss.item():.6f} oracle {loss_o.item():.6f}")
for k in ("lstm.weight_ih_l0", "lstm.weight_hh_l0", "lstm.bias_ih_l0", "fc_layers.0.weight", "fc_layers.9.weight", "fc_out.weight",
          "feature_extractor.bn4.weight", "feature_extractor.block5.rep.1.pointwise.weight", "feature_extractor.conv1.weight"):
    print(f"     {k:50s} ours {rel(dict(m.named_parameters())[k].grad, fo[k].grad):.3e}")
