import sys, torch
sys.path.insert(0, ".")
from oracle import xception_oracle as O
from multimodal_deepfake_detection_b200 import Block
torch.backends.cuda.matmul.allow_tf32 = False; torch.backends.cudnn.allow_tf32 = False
DEV="cuda"
def rel(a,b): return ((a.float()-b.float()).norm()/(b.float().norm()+1e-20)).item()
def leaf(sd): return {k:(v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone()) for k,v in sd.items()}
for cfg in [(728,728,3,1,True,True,19),(64,128,2,2,False,True,37)]:
    cin,cout,reps,stride,swr,gf,hw=cfg
    torch.manual_seed(5)
    blk=Block(cin,cout,reps,stride,start_with_relu=swr,grow_first=gf).to(DEV).eval()
    with torch.no_grad():
        for m_ in blk.modules():
            if isinstance(m_, torch.nn.BatchNorm2d):
                m_.running_mean.normal_(0,0.1); m_.running_var.uniform_(0.5,1.5); m_.weight.uniform_(0.5,1.5); m_.bias.normal_(0,0.1)
    x=(torch.randn(6,cin,hw,hw,device=DEV)*0.7).to(torch.bfloat16).float()
    sd={"b."+k:v.clone() for k,v in blk.state_dict().items()}
    res={}
    dout=None
    for tag in ("fp32","bf16"):
        lv=leaf(sd); xr=x.clone().requires_grad_(True)
        if tag=="bf16":
            with torch.autocast("cuda",dtype=torch.bfloat16): o=O.block_forward(lv,"b",("b",)+cfg[:6],xr,False,{}).float()
        else: o=O.block_forward(lv,"b",("b",)+cfg[:6],xr,False,{})
        if dout is None: dout=torch.randn_like(o)
        o.backward(dout); res[tag]={k:lv["b."+k].grad for k,_ in blk.named_parameters()}
    xo=x.clone().requires_grad_(True); out=blk(xo); out.backward(dout)
    for k,p in blk.named_parameters():
        print(f"{cfg[:2]} {k:28s} ours {rel(p.grad,res['fp32'][k]):.3e}  bf16-autocast {rel(res['bf16'][k],res['fp32'][k]):.3e}  |g| {res['fp32'][k].norm().item():.3e}")
