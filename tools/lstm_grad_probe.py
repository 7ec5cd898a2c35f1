"""LSTM + head gradients of XceptionLSTMA(512 / 128): this path vs the fp32 oracle vs the oracle fed bf16-rounded weights and
features, cluster vs single-CTA kernels (dev probe behind tests/test_models_gpu.py::test_wide_lstm_and_head_train_step_vs_oracle)."""
import os, sys, warnings
sys.path.insert(0, ".")
import torch, torch.nn.functional as F
from oracle import xception_oracle as O
from multimodal_deepfake_detection_b200 import XceptionLSTMA, BCELoss
DEV="cuda"
def rel(a,b): return ((a.float()-b.float()).norm()/(b.float().norm()+1e-20)).item()
def run(hidden,B,T):
    feat_sd = O.synth_state_dict(1234, num_classes=None, bn_jitter=0.1)
    full = {"feature_extractor." + k: v for k, v in feat_sd.items()}
    full.update(O.synth_lstm_head_state_dict(77, hidden))
    full = {k: v.to(DEV) for k, v in full.items()}
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        m = XceptionLSTMA(hidden).to(DEV)
    m.load_state_dict(full); m.train()
    for mod in m.modules():
        if isinstance(mod, torch.nn.Dropout): mod.eval()
    g = torch.Generator().manual_seed(31)
    feats = (torch.rand(B, T, 2048, generator=g) * 0.6).to(DEV)
    y = torch.randint(0, 2, (B, 1), generator=g).float().to(DEV)
    fo = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in full.items()}
    out_o,_,_ = O.lstm_forward(fo, feats); prob_o = O.head_forward(fo, out_o[:,-1]); F.binary_cross_entropy(prob_o,y).backward()
    # bf16-rounded-input oracle
    fb = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k, v in full.items()}
    with torch.no_grad():
        fb["lstm.weight_ih_l0"].copy_(fb["lstm.weight_ih_l0"].bfloat16().float()); fb["lstm.weight_hh_l0"].copy_(fb["lstm.weight_hh_l0"].bfloat16().float())
    out_b,_,_ = O.lstm_forward(fb, feats.bfloat16().float()); prob_b = O.head_forward(fb, out_b[:,-1]); F.binary_cross_entropy(prob_b,y).backward()
    prob = m(feats); BCELoss()(prob,y).backward()
    P = dict(m.named_parameters())
    print(hidden, B, T, os.environ.get("XCP_LSTM_NO_CLUSTER"), "prob err", (prob-prob_o).abs().max().item())
    for k in P:
        if k.startswith("feature_extractor"): continue
        print("   %-28s ours-vs-fp32 %.2e   bf16oracle-vs-fp32 %.2e  ours-vs-bf16oracle %.2e  |g| %.2e" % (k, rel(P[k].grad, fo[k].grad), rel(fb[k].grad, fo[k].grad), rel(P[k].grad, fb[k].grad), fo[k].grad.norm().item()))
run(512,5,24)
os.environ["XCP_LSTM_NO_CLUSTER"]="1"
run(512,5,24)
run(128,5,24)
