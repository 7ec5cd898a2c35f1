"""Quick GPU timing probe: XceptionLSTMV train step pieces on synthetic clips (CUDA events)."""
import sys, time
import torch
sys.path.insert(0, ".")
from multimodal_deepfake_detection_b200 import XceptionLSTMV
import torch.nn.functional as F
dev = "cuda"
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
H = int(sys.argv[2]) if len(sys.argv) > 2 else 299
frozen = (sys.argv[3] == "frozen") if len(sys.argv) > 3 else False
torch.manual_seed(0)
m = XceptionLSTMV(128).to(dev)
m.train()
for p in m.feature_extractor.parameters():
    p.requires_grad = not frozen
clips = torch.rand(B, 16, 3, H, H, device=dev)
y = torch.randint(0, 2, (B, 1), device=dev).float()
opt = torch.optim.Adam([p for p in m.parameters()], lr=1e-5, weight_decay=1e-4)
def step():
    opt.zero_grad(set_to_none=True)
    feats = m.extract_features(clips, torch.device(dev))
    prob = m(feats)
    loss = F.binary_cross_entropy(prob, y)
    loss.backward()
    opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = 5
t0 = time.time(); e0.record()
for _ in range(n): l = step()
e1.record(); torch.cuda.synchronize(); t1 = time.time()
ms = e0.elapsed_time(e1) / n
print(f"B={B} H={H} frozen={frozen}: {ms:.2f} ms/step (wall {(t1-t0)/n*1e3:.2f}) -> {B/ms*1e3:.1f} clips/s, {B*16/ms*1e3:.0f} frames/s, loss {l.item():.4f}, mem {torch.cuda.max_memory_allocated()/2**30:.1f} GiB")
# fwd only
with torch.no_grad():
    for _ in range(2): m.extract_features(clips, torch.device(dev))
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): m.extract_features(clips, torch.device(dev))
    e1.record(); torch.cuda.synchronize()
print(f"   extract_features no_grad: {e0.elapsed_time(e1)/n:.2f} ms")
if len(sys.argv) > 4:
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step(); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=35, max_name_column_width=70))
