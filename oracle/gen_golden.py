"""ORACLE SUPPORT — generates tests/golden/*.npz by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_shim.py) on seeded synthetic inputs.

Run here (the container that has /root/reference):  python oracle/gen_golden.py
The GPU box has no reference tree; it only reads the committed vectors.

Weights come from oracle.xception_oracle.synth_state_dict / synth_lstm_head_state_dict
(seeded, construction-order independent) and are pushed into the reference modules with
load_state_dict(strict=True), so anyone can rebuild exactly the same weights.
"""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))

from oracle import ref_shim  # noqa: E402
from oracle import xception_oracle as O  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")


def _ref_classes_from_script(path, names):
    """Pull class definitions (ArcFaceHead, CBFocalLoss) out of a reference *script* whose
    top-level imports cannot be satisfied (missing Models.AUFaceModel) by
    compiling just those ClassDef nodes.  Executed in place; nothing is copied to the repo."""
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name in names]
    mod = ast.Module(body=body, type_ignores=[])
    ns = {"torch": torch, "nn": nn, "F": F, "np": np}
    exec(compile(mod, path, "exec"), ns)
    return [ns[n] for n in names]


def main():
    torch.set_num_threads(8)
    ref = ref_shim.load()
    os.makedirs(OUT, exist_ok=True)
    g = {}

    # ---- case A: bare Xception(num_classes=2), eval + train, 75x75 and one 299x299 frame
    sd = O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1)
    m = ref.Xception(num_classes=2)
    m.load_state_dict(sd, strict=True)
    gen = torch.Generator().manual_seed(0)
    x75 = torch.rand(2, 3, 75, 75, generator=gen)
    x299 = torch.rand(1, 3, 299, 299, generator=gen)
    m.eval()
    with torch.no_grad():
        g["A_eval_logits_75"] = m(x75).numpy()
        fc = m.fc
        m.fc = nn.Identity()
        g["A_eval_feat_75"] = m(x75).numpy()
        g["A_eval_feat_299"] = m(x299).numpy()
        m.fc = fc
    m.train()
    labels = torch.tensor([0, 1])
    logits = m(x75)
    loss = F.cross_entropy(logits, labels)
    loss.backward()
    g["A_train_logits_75"] = logits.detach().numpy()
    g["A_train_loss_75"] = loss.detach().numpy()
    names, norms = [], []
    for k, p in m.named_parameters():
        names.append(k)
        norms.append(p.grad.norm().item())
    g["A_grad_names"] = np.array(names)
    g["A_grad_norms"] = np.array(norms, dtype=np.float64)
    for k in ("conv1.weight", "bn1.weight", "bn1.bias", "block1.rep.0.conv1.weight", "block4.rep.1.conv1.weight",
              "block1.skipbn.weight", "bn4.bias", "fc.weight"):
        g["A_grad::" + k] = dict(m.named_parameters())[k].grad.numpy()
    new_sd = m.state_dict()
    for k in ("bn1", "block1.skipbn", "block4.rep.2", "block12.rep.5", "bn4"):
        g["A_rm::" + k] = new_sd[k + ".running_mean"].numpy()
        g["A_rv::" + k] = new_sd[k + ".running_var"].numpy()
        g["A_nbt::" + k] = new_sd[k + ".num_batches_tracked"].numpy()

    # ---- case B: XceptionLSTMV(32): eval probs; train-mode BN with dropout disabled + grads
    feat_sd = O.synth_state_dict(1234, num_classes=None, bn_jitter=0.1)
    lh = O.synth_lstm_head_state_dict(77, hidden=32)
    full = {"feature_extractor." + k: v for k, v in feat_sd.items()}
    full.update(lh)
    mv = ref.XceptionLSTMV(hidden_dim=32)
    mv.load_state_dict(full, strict=True)
    clips = torch.rand(2, 3, 3, 75, 75, generator=gen)
    mv.eval()
    with torch.no_grad():
        feats = mv.extract_features(clips, torch.device("cpu"))
        g["B_eval_feats"] = feats.numpy()
        g["B_eval_probs"] = mv(feats).numpy()
        g["B_eval_lstm_out"] = mv.lstm(feats)[0].numpy()
    mv.train()
    for mod in mv.modules():
        if isinstance(mod, nn.Dropout):
            mod.eval()
    for p in mv.feature_extractor.parameters():      # unfreeze, train_visual.py:555-556
        p.requires_grad = True
    feats = mv.extract_features(clips, torch.device("cpu"))
    probs = mv(feats)
    y = torch.tensor([[1.0], [0.0]])
    loss = nn.BCELoss()(probs, y)                     # train_audio.py:20,39
    loss.backward()
    g["B_train_probs"] = probs.detach().numpy()
    g["B_train_loss"] = loss.detach().numpy()
    names, norms = [], []
    for k, p in mv.named_parameters():
        names.append(k)
        norms.append(p.grad.norm().item())
    g["B_grad_names"] = np.array(names)
    g["B_grad_norms"] = np.array(norms, dtype=np.float64)
    for k in ("lstm.weight_hh_l0", "lstm.bias_ih_l0", "fc_out.weight", "fc_layers.0.bias"):
        g["B_grad::" + k] = dict(mv.named_parameters())[k].grad.numpy()

    # ---- case C: XceptionLSTMA(32) on MFCC-like (B,T,3,13)
    ma = ref.XceptionLSTMA(hidden_dim=32)
    ma.load_state_dict(full, strict=True)
    audio1 = torch.randn(2, 4, 1, 13, generator=gen) * 20.0
    audio = audio1.repeat(1, 1, 3, 1).contiguous()      # audio_dataloader.py:25-26
    ma.eval()
    with torch.no_grad():
        fa = ma.extract_features(audio, torch.device("cpu"))
        g["C_eval_feats"] = fa.numpy()
        g["C_eval_probs"] = ma(fa).numpy()
    g["C_audio"] = audio.numpy()

    # ---- case D: ArcFace + CB-focal + fusion regularisers (train_au_face.py:423-458,659-674)
    ArcFaceHead, CBFocalLoss = _ref_classes_from_script(os.path.join(ref_shim.REF_DIR, "train_au_face.py"),
                                                        ["ArcFaceHead", "CBFocalLoss"])
    (ArcFaceV,) = _ref_classes_from_script(os.path.join(ref_shim.REF_DIR, "train_visual.py"), ["ArcFaceHead"])
    B, T, D = 4, 5, 24
    v_tok = torch.randn(B, T, D, generator=gen, requires_grad=True)
    a_tok = torch.randn(B, T, D, generator=gen, requires_grad=True)
    lab = torch.tensor([0, 1, 1, 0])
    embed = nn.Sequential(nn.Linear(2 * D, 256), nn.ReLU(inplace=True), nn.Dropout(0.2), nn.Linear(256, 128))
    ge = torch.Generator().manual_seed(5)
    with torch.no_grad():
        for p in embed.parameters():
            p.copy_(torch.randn(p.shape, generator=ge) * 0.1)
    embed.eval()
    arc = ArcFaceHead(128, 2, s=30.0, m=0.30)
    with torch.no_grad():
        arc.weight.copy_(torch.randn(2, 128, generator=ge) * 0.2)
    cb = CBFocalLoss([500, 10000], beta=0.9999, gamma=2.0)
    v_pool, a_pool = v_tok.mean(1), a_tok.mean(1)
    emb = embed(torch.cat([v_pool, a_pool], 1))
    logits_arc = arc(emb, lab)
    loss_cls = cb(logits_arc, lab)
    loss_align = F.mse_loss(v_pool, a_pool)
    lt = 0.5 * ((v_tok[:, 1:] - v_tok[:, :-1]).pow(2).mean() + (a_tok[:, 1:] - a_tok[:, :-1]).pow(2).mean())
    loss = loss_cls + 0.2 * loss_align + 0.1 * lt
    loss.backward()
    g["D_v_tok"] = v_tok.detach().numpy(); g["D_a_tok"] = a_tok.detach().numpy()
    g["D_labels"] = lab.numpy()
    for k, p in embed.state_dict().items():
        g["D_embed::" + k] = p.numpy()
    g["D_arc_w"] = arc.weight.detach().numpy()
    g["D_class_weights"] = cb.class_weights.numpy()
    g["D_logits"] = logits_arc.detach().numpy()
    g["D_loss"] = loss.detach().numpy()
    g["D_grad_v_tok"] = v_tok.grad.numpy(); g["D_grad_a_tok"] = a_tok.grad.numpy()
    g["D_grad_arc_w"] = arc.weight.grad.numpy()
    g["D_grad_embed0_w"] = embed[0].weight.grad.numpy()
    # visual ArcFace (s=30, m=0.5) + CE, train_visual.py:455-474,532
    arcv = ArcFaceV(32, 2, s=30.0, m=0.5)
    with torch.no_grad():
        arcv.weight.copy_(torch.randn(2, 32, generator=ge) * 0.2)
    e = torch.randn(4, 32, generator=ge, requires_grad=True)
    lg = arcv(e, lab)
    l2 = F.cross_entropy(lg, lab)
    l2.backward()
    g["D2_emb"] = e.detach().numpy(); g["D2_w"] = arcv.weight.detach().numpy()
    g["D2_logits"] = lg.detach().numpy(); g["D2_loss"] = l2.detach().numpy()
    g["D2_grad_emb"] = e.grad.numpy(); g["D2_grad_w"] = arcv.weight.grad.numpy()
    with torch.no_grad():
        g["D2_logits_nolabel"] = arcv(e).numpy()

    # ---- inputs so that tests never depend on torch's RNG stream staying stable
    g["x75"] = x75.numpy(); g["x299"] = x299.numpy(); g["clips"] = clips.numpy()
    np.savez_compressed(os.path.join(OUT, "reference_golden.npz"), **g)
    print("wrote", os.path.join(OUT, "reference_golden.npz"), "keys:", len(g))


if __name__ == "__main__":
    main()
