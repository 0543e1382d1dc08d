"""TEST INFRASTRUCTURE.  Measures the bf16 noise floor of the REFERENCE ITSELF: the unmodified reference modules run
under torch.autocast(bfloat16) against their own fp32 run (same seed, eval mode).  The numbers calibrate the
tolerances stated in tests/test_gpu_models.py.  Build-container only (needs /root/reference).

Output recorded on 2026-10-17 (torch 2.11, 8 CPU threads):
  small D=64  : x.grad rel-L2 8.1e-2, probs max-abs 4.3e-3, parameter grads rel-L2 7.2e-2 .. 1.7e-1
  C2   D=2048 : x.grad rel-L2 7.7e-2, probs max-abs 6.2e-3, parameter grads rel-L2 6.9e-2 .. 1.5e-1 (head 1.5e-1)
  C4   D=1024, L=81, 12 windows: x.grad 1.2e-1, probs 8.1e-3, parameter grads up to 2.2e-1 (head.classifier.0.bias
                2.2e-1, max element error 0.61 x max|ref|)
"""
import sys, torch, types
sys.path.insert(0,'/root/reference'); sys.path.insert(0,'/root/reference/Train')
import importlib
for n in ("h5py","matplotlib","matplotlib.pyplot"):
    sys.modules.setdefault(n, types.ModuleType(n))
from models.Encoder import Encoder
from models.Classifier import Classifier
ltn = importlib.import_module("temporal_transformer_shanghaitech")
torch.set_num_threads(8)
def run(kw,B,P,T,N,autocast):
    D=kw['d_model']
    torch.manual_seed(0)
    enc=Encoder(**kw).eval(); cls=Classifier(D,0.6).eval()
    g=torch.Generator().manual_seed(1)
    x=torch.randn(2*B*P,T*N,D,generator=g).abs().requires_grad_(True)
    m=(torch.rand(B,P,generator=g)>0.9).float()
    labs=torch.cat([torch.stack([torch.ones(B,P),torch.zeros(B,P)],-1), torch.stack([1-m,m],-1)],0).view(2*B*P,2)
    args=types.SimpleNamespace(batch_size=B,part_num=P,lambda_1=0.01)
    ctx = torch.autocast('cpu',dtype=torch.bfloat16) if autocast else torch.autocast('cpu',enabled=False)
    with ctx:
        out=enc(x)
        probs=cls(out[:,0,:].float().view(2*B,P,D)).view(2*B*P,-1).float()
    ce=ltn.get_CE_loss(args,probs,labs); mil,_,_=ltn.get_MIL_loss(args,probs[:,1])
    (mil+0.8*ce).backward()
    grads={k:p.grad.clone() for k,p in enc.named_parameters() if p.grad is not None}
    grads.update({'head.'+k:p.grad.clone() for k,p in cls.named_parameters() if p.grad is not None})
    return x.grad.clone(), grads, probs.detach()
for name,(kw,B,P,T,N) in {
 "small D=64": (dict(n_layers=2,n_head=2,d_k=64,d_v=64,d_model=64,d_inner=128,MHA_layerNorm=True,FFN_layerNorm=True,weight_init=False,relative_pe=True,window_size=4,window_depth=3),2,4,3,16),
 "C2 D=2048": (dict(n_layers=3,n_head=8,d_k=256,d_v=256,d_model=2048,d_inner=4096,MHA_layerNorm=True,FFN_layerNorm=True,weight_init=False,relative_pe=True,window_size=4,window_depth=3),2,4,3,16),
 "C4 D=1024 L=81 (12 windows)": (dict(n_layers=3,n_head=8,d_k=256,d_v=256,d_model=1024,d_inner=4096,MHA_layerNorm=True,FFN_layerNorm=True,weight_init=False,relative_pe=True,window_size=4,window_depth=5),2,3,5,16)}.items():
    xg0,g0,p0=run(kw,B,P,T,N,False)
    xg1,g1,p1=run(kw,B,P,T,N,True)
    rl2=lambda a,b:((a-b).double().norm()/b.double().norm()).item()
    print(name,"autocast vs fp32: x.grad rel_l2 %.3e"%rl2(xg1,xg0),"probs max %.2e"%(p1-p0).abs().max().item())
    for k in list(g0)[:1]+["layer_stack.1.pos_ffn.w_1.weight","layer_stack.0.slf_attn.fc.weight"]+[k for k in g0 if k.startswith("head.")]:
        d=(g1[k]-g0[k]).abs(); sc=g0[k].abs().max().item()
        print("   ",k,"rel_l2 %.3e  max %.3e x max|ref|"%(rl2(g1[k],g0[k]), d.max().item()/sc))
