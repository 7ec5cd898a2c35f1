"""The oracle (oracle/xception_oracle.py) against vectors produced by the UNMODIFIED
reference (oracle/gen_golden.py) and, when /root/reference is present, the live reference."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim
from oracle import xception_oracle as O


def _t(a):
    return torch.from_numpy(np.asarray(a))


def test_eval_logits_and_features(golden):
    sd = O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1)
    x = _t(golden["x75"])
    with torch.no_grad():
        logits = O.xception_logits(sd, x, training=False)
        feat = O.xception_features(sd, x, training=False)
    assert torch.allclose(logits, _t(golden["A_eval_logits_75"]), rtol=1e-5, atol=1e-6)
    assert torch.allclose(feat, _t(golden["A_eval_feat_75"]), rtol=1e-5, atol=1e-6)


def test_eval_299(golden):
    sd = O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1)
    with torch.no_grad():
        feat = O.xception_features(sd, _t(golden["x299"]), training=False)
    assert torch.allclose(feat, _t(golden["A_eval_feat_299"]), rtol=1e-4, atol=1e-5)


def test_train_logits_grads_and_running_stats(golden):
    sd = O.synth_state_dict(1234, num_classes=2, bn_jitter=0.1)
    names = [str(n) for n in golden["A_grad_names"]]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    new_stats = {}
    logits = O.xception_logits(sd, _t(golden["x75"]), training=True, new_stats=new_stats)
    loss = F.cross_entropy(logits, torch.tensor([0, 1]))
    loss.backward()
    assert torch.allclose(logits.detach(), _t(golden["A_train_logits_75"]), rtol=1e-4, atol=1e-5)
    assert abs(loss.item() - float(golden["A_train_loss_75"])) < 1e-5
    norms = np.array([sd[n].grad.norm().item() for n in names])
    ref = golden["A_grad_norms"]
    assert np.all(np.abs(norms - ref) <= 1e-2 * np.abs(ref) + 1e-7)   # fp32 re-association noise
    for k in ("conv1.weight", "bn1.weight", "block4.rep.1.conv1.weight", "fc.weight"):
        gref = _t(golden["A_grad::" + k])
        assert (sd[k].grad - gref).norm() <= 1e-2 * gref.norm()   # tiny-batch train-mode BN amplifies fp32 re-association noise
    for k in ("bn1", "block1.skipbn", "block4.rep.2", "block12.rep.5", "bn4"):
        assert torch.allclose(new_stats[k + ".running_mean"], _t(golden["A_rm::" + k]), rtol=1e-4, atol=1e-6)
        assert torch.allclose(new_stats[k + ".running_var"], _t(golden["A_rv::" + k]), rtol=1e-4, atol=1e-6)
        assert int(new_stats[k + ".num_batches_tracked"]) == int(golden["A_nbt::" + k])


def _lstm_sd():
    feat_sd = O.synth_state_dict(1234, num_classes=None, bn_jitter=0.1)
    full = {"feature_extractor." + k: v for k, v in feat_sd.items()}
    full.update(O.synth_lstm_head_state_dict(77, hidden=32))
    return full


def test_lstmv_eval_and_train(golden):
    sd = _lstm_sd()
    clips = _t(golden["clips"])
    with torch.no_grad():
        probs = O.xception_lstm_forward(sd, clips, training=False)
        b, t = clips.shape[:2]
        f = O.xception_features(sd, clips.reshape(b * t, *clips.shape[2:]), False, None, "feature_extractor.")
        out, _, _ = O.lstm_forward(sd, f.view(b, t, -1))
    assert torch.allclose(f.view(b, t, -1), _t(golden["B_eval_feats"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(out, _t(golden["B_eval_lstm_out"]), rtol=1e-4, atol=1e-5)
    assert torch.allclose(probs, _t(golden["B_eval_probs"]), rtol=1e-4, atol=1e-6)
    names = [str(n) for n in golden["B_grad_names"]]
    for n in names:
        sd[n] = sd[n].clone().requires_grad_(True)
    probs = O.xception_lstm_forward(sd, clips, training=True, new_stats={})
    loss = F.binary_cross_entropy(probs, torch.tensor([[1.0], [0.0]]))
    loss.backward()
    assert torch.allclose(probs.detach(), _t(golden["B_train_probs"]), rtol=1e-4, atol=1e-6)
    norms = np.array([sd[n].grad.norm().item() for n in names])
    ref = golden["B_grad_norms"]
    assert np.all(np.abs(norms - ref) <= 1e-2 * np.abs(ref) + 1e-8)
    for k in ("lstm.weight_hh_l0", "lstm.bias_ih_l0", "fc_out.weight", "fc_layers.0.bias"):
        gref = _t(golden["B_grad::" + k])
        assert (sd[k].grad - gref).norm() <= 1e-2 * gref.norm()   # tiny-batch train-mode BN amplifies fp32 re-association noise


def test_lstma_eval(golden):
    sd = _lstm_sd()
    with torch.no_grad():
        probs = O.xception_lstm_forward(sd, _t(golden["C_audio"]), training=False, audio=True)
    assert torch.allclose(probs, _t(golden["C_eval_probs"]), rtol=1e-4, atol=1e-6)


def test_fusion_head_and_arcface(golden):
    v = _t(golden["D_v_tok"]).requires_grad_(True)
    a = _t(golden["D_a_tok"]).requires_grad_(True)
    lab = _t(golden["D_labels"])
    esd = {k: _t(golden["D_embed::" + k]).clone().requires_grad_(True)
           for k in ("0.weight", "0.bias", "3.weight", "3.bias")}
    arc_w = _t(golden["D_arc_w"]).clone().requires_grad_(True)
    cw = O.cb_focal_weights([500, 10000])
    assert torch.allclose(cw, _t(golden["D_class_weights"]), rtol=1e-6)
    loss, logits = O.fusion_head_loss(esd, arc_w, v, a, lab, cw)
    loss.backward()
    assert torch.allclose(logits.detach(), _t(golden["D_logits"]), rtol=1e-5, atol=1e-5)
    assert abs(loss.item() - float(golden["D_loss"])) < 1e-5
    assert torch.allclose(v.grad, _t(golden["D_grad_v_tok"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(a.grad, _t(golden["D_grad_a_tok"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(arc_w.grad, _t(golden["D_grad_arc_w"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(esd["0.weight"].grad, _t(golden["D_grad_embed0_w"]), rtol=1e-4, atol=1e-7)
    e = _t(golden["D2_emb"]).requires_grad_(True)
    w = _t(golden["D2_w"]).clone().requires_grad_(True)
    lg = O.arcface_logits(w, e, lab, 30.0, 0.5)
    l2 = F.cross_entropy(lg, lab)
    l2.backward()
    assert torch.allclose(lg.detach(), _t(golden["D2_logits"]), rtol=1e-5, atol=1e-5)
    assert torch.allclose(e.grad, _t(golden["D2_grad_emb"]), rtol=1e-4, atol=1e-7)
    assert torch.allclose(w.grad, _t(golden["D2_grad_w"]), rtol=1e-4, atol=1e-7)
    with torch.no_grad():
        assert torch.allclose(O.arcface_logits(w, e, None, 30.0, 0.5), _t(golden["D2_logits_nolabel"]), rtol=1e-5, atol=1e-5)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree not mounted (GPU box)")
def test_oracle_against_live_reference_and_schema():
    ref = ref_shim.load()
    torch.manual_seed(3)
    m = ref.Xception(num_classes=2)
    sd = m.state_dict()
    # schema restatement: same keys, order, shapes as the real constructor
    shapes = O.xception_param_shapes(2)
    assert [k for k, _, _ in shapes] == list(sd.keys())
    assert all(tuple(sd[k].shape) == tuple(s) for k, s, _ in shapes)
    assert len(sd) == 276
    x = torch.rand(2, 3, 71, 71)
    m.eval()
    with torch.no_grad():
        assert torch.allclose(m(x), O.xception_logits(sd, x, False), rtol=1e-5, atol=1e-6)
    mv = ref.XceptionLSTMV(hidden_dim=16)
    assert len(mv.state_dict()) == 288
