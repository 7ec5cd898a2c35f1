"""ORACLE — TEST INFRASTRUCTURE ONLY.  Not product code.

A plain fp32 PyTorch *functional* restatement of the reference hot path
(Tonmoy1321/Multimodal-DeepFake-Detection): Xception backbone, the LSTM over
per-frame features, the MLP/sigmoid head, ArcFace head, CB-focal loss and the
audio-face fusion loss.  Every function works on a flat ``state_dict``
(``{key: tensor}``) with the reference's key names, so the same seeded weights
can be fed to the reference modules, to this oracle and to the sm_100a path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import this file.  The product
package never does (``tests/test_cabi_and_host_cpu.py::test_product_never_imports_the_oracle`` enforces that).

Parity pinning: the reference ships NO golden vectors or tests,
so this oracle is pinned against the reference *itself*: ``oracle/gen_golden.py``
imports the untouched reference modules from ``/root/reference`` (through the
package shim in ``oracle/ref_shim.py``), runs them on seeded inputs and commits
the outputs under ``tests/golden/``; ``tests/test_oracle_vs_golden.py`` checks
this restatement against those vectors (and against the live reference when
``/root/reference`` is present).

Citations are ``file:line`` in ``/root/reference``.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor
SD = Dict[str, Tensor]

BN_EPS = 1e-5        # nn.BatchNorm2d default, Xception.py:56,67,73,78,119,123,143,147
BN_MOMENTUM = 0.1

# (name, in, out, reps, stride, start_with_relu, grow_first)   Xception.py:126-140
BLOCKS = [
    ("block1", 64, 128, 2, 2, False, True),
    ("block2", 128, 256, 2, 2, True, True),
    ("block3", 256, 728, 2, 2, True, True),
] + [("block%d" % i, 728, 728, 3, 1, True, True) for i in range(4, 12)] + [
    ("block12", 728, 1024, 2, 2, True, False),
]


def block_layout(cin: int, cout: int, reps: int, start_with_relu: bool, grow_first: bool):
    """Return [(rep_index_of_sepconv, rep_index_of_bn, c_in, c_out, relu_before)].

    Restates the list building in Block.__init__ (Xception.py:61-87): each unit
    is [ReLU, SeparableConv2d, BatchNorm2d]; the very first ReLU is dropped when
    ``start_with_relu`` is False (which shifts every Sequential index by -1).
    """
    units = []
    filters = cin
    if grow_first:
        units.append((cin, cout))
        filters = cout
    for _ in range(reps - 1):
        units.append((filters, filters))
    if not grow_first:
        units.append((cin, cout))
    out = []
    idx = 0
    for u, (ci, co) in enumerate(units):
        relu = True
        if u == 0 and not start_with_relu:
            relu = False
        else:
            idx += 1  # the ReLU occupies a Sequential slot
        out.append((idx, idx + 1, ci, co, relu))
        idx += 2
    return out


def _bn(sd: SD, key: str, x: Tensor, training: bool, new_stats: Optional[SD]) -> Tensor:
    """nn.BatchNorm2d forward: biased var for normalisation,
    unbiased var into running_var, momentum 0.1."""
    w, b = sd[key + ".weight"], sd[key + ".bias"]
    rm, rv = sd[key + ".running_mean"], sd[key + ".running_var"]
    if training:
        mean = x.mean(dim=(0, 2, 3))
        var = x.var(dim=(0, 2, 3), unbiased=False)
        if new_stats is not None:
            n = x.numel() // x.shape[1]
            new_stats[key + ".running_mean"] = (1 - BN_MOMENTUM) * rm + BN_MOMENTUM * mean.detach()
            new_stats[key + ".running_var"] = (1 - BN_MOMENTUM) * rv + BN_MOMENTUM * var.detach() * (n / max(n - 1, 1))
            new_stats[key + ".num_batches_tracked"] = sd[key + ".num_batches_tracked"] + 1
    else:
        mean, var = rm, rv
    scale = w * torch.rsqrt(var + BN_EPS)
    shift = b - mean * scale
    return x * scale[None, :, None, None] + shift[None, :, None, None]


def sepconv(sd: SD, key: str, x: Tensor) -> Tensor:
    """SeparableConv2d.forward (Xception.py:44-47): depthwise 3x3 s1 p1 then 1x1."""
    c = x.shape[1]
    x = F.conv2d(x, sd[key + ".conv1.weight"], None, 1, 1, 1, groups=c)
    return F.conv2d(x, sd[key + ".pointwise.weight"])


def block_forward(sd: SD, name: str, cfg, inp: Tensor, training: bool, new_stats: Optional[SD]) -> Tensor:
    """Block.forward (Xception.py:89-99)."""
    _, cin, cout, reps, stride, swr, gf = cfg
    x = inp
    for (i_sep, i_bn, _ci, _co, relu) in block_layout(cin, cout, reps, swr, gf):
        if relu:
            x = F.relu(x)          # first ReLU of rep is out-of-place (Xception.py:83)
        x = sepconv(sd, f"{name}.rep.{i_sep}", x)
        x = _bn(sd, f"{name}.rep.{i_bn}", x, training, new_stats)
    if stride != 1:
        x = F.max_pool2d(x, 3, stride, 1)      # Xception.py:86
    if cin != cout or stride != 1:             # Xception.py:54-58, 92-94
        skip = F.conv2d(inp, sd[f"{name}.skip.weight"], None, stride)
        skip = _bn(sd, f"{name}.skipbn", skip, training, new_stats)
    else:
        skip = inp
    return x + skip


def xception_features(sd: SD, x: Tensor, training: bool = False, new_stats: Optional[SD] = None,
                      prefix: str = "") -> Tensor:
    """Xception.forward up to (and including) global-average-pool (Xception.py:167-198)."""
    p = {k[len(prefix):]: v for k, v in sd.items() if k.startswith(prefix)} if prefix else sd
    x = F.conv2d(x, p["conv1.weight"], None, 2)                       # :168
    x = F.relu(_bn(p, "bn1", x, training, new_stats))                 # :169-170
    x = F.conv2d(x, p["conv2.weight"])                                # :172
    x = F.relu(_bn(p, "bn2", x, training, new_stats))                 # :173-174
    for cfg in BLOCKS:                                                # :176-187
        x = block_forward(p, cfg[0], cfg, x, training, new_stats)
    x = sepconv(p, "conv3", x)                                        # :189
    x = F.relu(_bn(p, "bn3", x, training, new_stats))                 # :190-191
    x = sepconv(p, "conv4", x)                                        # :193
    x = F.relu(_bn(p, "bn4", x, training, new_stats))                 # :194-195
    x = F.adaptive_avg_pool2d(x, (1, 1)).flatten(1)                   # :197-198
    if prefix and new_stats is not None:
        for k in list(new_stats.keys()):
            if not k.startswith(prefix):
                new_stats[prefix + k] = new_stats.pop(k)
    return x


def xception_logits(sd: SD, x: Tensor, training: bool = False, new_stats: Optional[SD] = None) -> Tensor:
    """Full Xception.forward incl. fc (Xception.py:199)."""
    f = xception_features(sd, x, training, new_stats)
    return F.linear(f, sd["fc.weight"], sd["fc.bias"])


def bilinear_13_to_64(audio: Tensor) -> Tensor:
    """XceptionLSTMA.extract_features reshape + interpolate (XceptionLSTMA.py:45-46)."""
    b, t, c, n = audio.shape
    frames = audio.reshape(b * t, c, n, 1)
    return F.interpolate(frames, size=(64, 64), mode="bilinear", align_corners=False)


def lstm_forward(sd: SD, feats: Tensor, prefix: str = "lstm.") -> Tuple[Tensor, Tensor, Tensor]:
    """nn.LSTM(2048,H,1,batch_first) with zero initial state (XceptionLSTMV.py:18-23,67).
    Gate order i,f,g,o; two bias vectors."""
    w_ih, w_hh = sd[prefix + "weight_ih_l0"], sd[prefix + "weight_hh_l0"]
    b = sd[prefix + "bias_ih_l0"] + sd[prefix + "bias_hh_l0"]
    bsz, t, _ = feats.shape
    hdim = w_hh.shape[1]
    h = feats.new_zeros(bsz, hdim)
    c = feats.new_zeros(bsz, hdim)
    outs = []
    for s in range(t):
        g = feats[:, s] @ w_ih.t() + h @ w_hh.t() + b
        i, f, gg, o = g.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.stack(outs, 1), h, c


def head_forward(sd: SD, h_last: Tensor, drop_masks: Optional[List[Tensor]] = None,
                 return_logit: bool = False) -> Tensor:
    """fc_layers + fc_out + sigmoid (XceptionLSTMV.py:25-44,68-70).  ``drop_masks``
    are optional pre-scaled keep masks (value 0 or 1/(1-p)) for the 4 Dropout(0.3)."""
    x = h_last
    for li, k in enumerate((0, 3, 6, 9)):
        x = F.relu(F.linear(x, sd[f"fc_layers.{k}.weight"], sd[f"fc_layers.{k}.bias"]))
        if drop_masks is not None:
            x = x * drop_masks[li]
    z = F.linear(x, sd["fc_out.weight"], sd["fc_out.bias"])
    return z if return_logit else torch.sigmoid(z)


def xception_lstm_forward(sd: SD, clips: Tensor, training: bool = False, new_stats: Optional[SD] = None,
                          audio: bool = False, drop_masks=None) -> Tensor:
    """XceptionLSTMV/A.extract_features + forward (XceptionLSTMV.py:46-70, XceptionLSTMA.py:39-59)."""
    if audio:
        b, t = clips.shape[:2]
        frames = bilinear_13_to_64(clips)
    else:
        b, t, c, h, w = clips.shape
        frames = clips.reshape(b * t, c, h, w)
    f = xception_features(sd, frames, training, new_stats, prefix="feature_extractor.")
    out, _, _ = lstm_forward(sd, f.view(b, t, -1))
    return head_forward(sd, out[:, -1], drop_masks)


def arcface_logits(weight: Tensor, feats: Tensor, labels: Optional[Tensor], s: float, m: float) -> Tensor:
    """ArcFaceHead.forward (train_visual.py:464-474 / train_au_face.py:432-442)."""
    x = F.normalize(feats)
    w = F.normalize(weight)
    cos = x @ w.t()
    if labels is None:
        return s * cos
    theta = torch.acos(cos.clamp(-1 + 1e-7, 1 - 1e-7))
    target = torch.cos(theta + m)
    one_hot = F.one_hot(labels, num_classes=weight.shape[0]).float()
    return s * (cos * (1 - one_hot) + target * one_hot)


def cb_focal_weights(samples_per_cls, beta: float = 0.9999) -> Tensor:
    """CBFocalLoss.__init__ (train_au_face.py:447-452)."""
    n = torch.tensor(samples_per_cls, dtype=torch.float64)
    eff = 1.0 - torch.pow(torch.tensor(beta, dtype=torch.float64), n)
    w = (1.0 - beta) / eff
    w = w / w.sum() * len(samples_per_cls)
    return w.float()


def cb_focal_loss(logits: Tensor, labels: Tensor, class_weights: Tensor, gamma: float = 2.0) -> Tensor:
    """CBFocalLoss.forward (train_au_face.py:455-458)."""
    ce = F.cross_entropy(logits, labels, reduction="none", weight=class_weights)
    pt = torch.exp(-ce)
    return ((1 - pt) ** gamma * ce).mean()


def fusion_head_loss(embed_sd: SD, arc_w: Tensor, v_tokens: Tensor, au_tokens: Tensor, labels: Tensor,
                     class_weights: Tensor, s: float = 30.0, m: float = 0.30, gamma: float = 2.0,
                     lambda_align: float = 0.2, lambda_temp: float = 0.1, drop_mask: Optional[Tensor] = None):
    """The fused region of train_au_face.py:659-674: mean-pool both token streams, concat,
    embed_head (Linear-ReLU-Dropout(0.2)-Linear), ArcFace, CB-focal, + align MSE + temporal
    smoothness.  Returns (loss, logits)."""
    v_pool = v_tokens.mean(1)
    au_pool = au_tokens.mean(1)
    pooled = torch.cat([v_pool, au_pool], dim=1)
    hdn = F.relu(F.linear(pooled, embed_sd["0.weight"], embed_sd["0.bias"]))
    if drop_mask is not None:
        hdn = hdn * drop_mask
    embed = F.linear(hdn, embed_sd["3.weight"], embed_sd["3.bias"])
    logits = arcface_logits(arc_w, embed, labels, s, m)
    loss_cls = cb_focal_loss(logits, labels, class_weights, gamma)
    loss_align = F.mse_loss(v_pool, au_pool)
    lt_v = (v_tokens[:, 1:] - v_tokens[:, :-1]).pow(2).mean() if v_tokens.size(1) > 1 else v_tokens.new_tensor(0.0)
    lt_a = (au_tokens[:, 1:] - au_tokens[:, :-1]).pow(2).mean() if au_tokens.size(1) > 1 else au_tokens.new_tensor(0.0)
    loss = loss_cls + lambda_align * loss_align + lambda_temp * 0.5 * (lt_v + lt_a)
    return loss, logits


def label_smoothing_bce_with_logits(logits: Tensor, targets: Tensor, smoothing: float = 0.1) -> Tensor:
    """LabelSmoothingBCEWithLogitsLoss (train_au_patch.py:203-211)."""
    t = targets * (1 - smoothing) + 0.5 * smoothing
    return F.binary_cross_entropy_with_logits(logits, t)


# ----------------------------------------------------------------------------
# Seeded weights without the reference: restates Xception.__init__'s init
# (Xception.py:155-160) so the GPU box (no /root/reference) can rebuild the same
# state_dict the golden vectors were made with.  Checked against the real
# constructor in tests/test_oracle_vs_reference.py.
# ----------------------------------------------------------------------------
def xception_param_shapes(num_classes: Optional[int] = 1000):
    """(key, shape, kind) in nn.Module registration order of the reference."""
    out = []

    def bn(k, c):
        out.extend([(k + ".weight", (c,), "bn_w"), (k + ".bias", (c,), "bn_b"),
                    (k + ".running_mean", (c,), "rm"), (k + ".running_var", (c,), "rv"),
                    (k + ".num_batches_tracked", (), "nbt")])

    out.append(("conv1.weight", (32, 3, 3, 3), "conv")); bn("bn1", 32)
    out.append(("conv2.weight", (64, 32, 3, 3), "conv")); bn("bn2", 64)
    for (name, cin, cout, reps, stride, swr, gf) in BLOCKS:
        if cin != cout or stride != 1:
            out.append((f"{name}.skip.weight", (cout, cin, 1, 1), "conv")); bn(f"{name}.skipbn", cout)
        for (i_sep, i_bn, ci, co, _r) in block_layout(cin, cout, reps, swr, gf):
            out.append((f"{name}.rep.{i_sep}.conv1.weight", (ci, 1, 3, 3), "conv"))
            out.append((f"{name}.rep.{i_sep}.pointwise.weight", (co, ci, 1, 1), "conv"))
            bn(f"{name}.rep.{i_bn}", co)
    out.append(("conv3.conv1.weight", (1024, 1, 3, 3), "conv"))
    out.append(("conv3.pointwise.weight", (1536, 1024, 1, 1), "conv")); bn("bn3", 1536)
    out.append(("conv4.conv1.weight", (1536, 1, 3, 3), "conv"))
    out.append(("conv4.pointwise.weight", (2048, 1536, 1, 1), "conv")); bn("bn4", 2048)
    if num_classes is not None:
        out.append(("fc.weight", (num_classes, 2048), "fc_w"))
        out.append(("fc.bias", (num_classes,), "fc_b"))
    return out


def synth_state_dict(seed: int, num_classes: Optional[int] = 2, bn_jitter: float = 0.0,
                     dtype=torch.float32) -> SD:
    """Deterministic *oracle-owned* weights: conv ~ N(0, sqrt(2/(kh*kw*out))) as in
    Xception.py:155-157, BN gamma=1/beta=0 (optionally jittered so that BN affine,
    running stats and negative gammas are exercised).  Independent of torch's module
    construction order, so it is reproducible anywhere."""
    g = torch.Generator().manual_seed(seed)
    sd: SD = {}
    for key, shape, kind in xception_param_shapes(num_classes):
        if kind == "conv":
            n = shape[2] * shape[3] * shape[0]
            sd[key] = torch.randn(shape, generator=g, dtype=dtype) * math.sqrt(2.0 / n)
        elif kind == "bn_w":
            sd[key] = torch.ones(shape, dtype=dtype)
            if bn_jitter:
                sd[key] += bn_jitter * torch.randn(shape, generator=g, dtype=dtype)
        elif kind == "bn_b":
            sd[key] = torch.zeros(shape, dtype=dtype)
            if bn_jitter:
                sd[key] += bn_jitter * torch.randn(shape, generator=g, dtype=dtype)
        elif kind == "rm":
            sd[key] = torch.zeros(shape, dtype=dtype)
            if bn_jitter:
                sd[key] += bn_jitter * torch.randn(shape, generator=g, dtype=dtype)
        elif kind == "rv":
            sd[key] = torch.ones(shape, dtype=dtype)
            if bn_jitter:
                sd[key] += bn_jitter * torch.rand(shape, generator=g, dtype=dtype)
        elif kind == "nbt":
            sd[key] = torch.zeros((), dtype=torch.long)
        elif kind == "fc_w":
            bound = 1.0 / math.sqrt(shape[1])
            sd[key] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
        elif kind == "fc_b":
            bound = 1.0 / math.sqrt(2048)
            sd[key] = (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * bound
    return sd


def synth_lstm_head_state_dict(seed: int, hidden: int, dtype=torch.float32) -> SD:
    """Seeded LSTM + MLP head weights with PyTorch's default U(-1/sqrt(fan), +) ranges
    (nn.LSTM / nn.Linear defaults used by XceptionLSTMV.py:18-43)."""
    g = torch.Generator().manual_seed(seed)

    def u(shape, fan):
        b = 1.0 / math.sqrt(fan)
        return (torch.rand(shape, generator=g, dtype=dtype) * 2 - 1) * b

    sd: SD = {
        "lstm.weight_ih_l0": u((4 * hidden, 2048), hidden),
        "lstm.weight_hh_l0": u((4 * hidden, hidden), hidden),
        "lstm.bias_ih_l0": u((4 * hidden,), hidden),
        "lstm.bias_hh_l0": u((4 * hidden,), hidden),
    }
    dims = [hidden, 1024, 1024, 1024, 1024]
    for li, k in enumerate((0, 3, 6, 9)):
        sd[f"fc_layers.{k}.weight"] = u((dims[li + 1], dims[li]), dims[li])
        sd[f"fc_layers.{k}.bias"] = u((dims[li + 1],), dims[li])
    sd["fc_out.weight"] = u((1, 1024), 1024)
    sd["fc_out.bias"] = u((1,), 1024)
    return sd
