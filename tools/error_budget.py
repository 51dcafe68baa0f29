"""Error budget of the fp16 tensor-core path (CPU, oracle only): emulate each fp16 rounding point of the kernels inside the
fp32 oracle of SurfPosNet and report the relative L2 error of the prediction it causes on its own and in combination.
This is the experiment behind the `precision` modes: weights are the dominant, systematic term, and
inside in_proj only the VALUE rows matter.

    python tools/error_budget.py
"""
import sys, math, torch
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch.nn.functional as F
from brepgen_b200.spec import denoiser_spec
from brepgen_b200.synth import synth_state_dict
from oracle import denoisers as O
torch.set_num_threads(8)
D,NH=768,12
def h(x,on): return x.half().float() if on else x
def run(sd, x, t, R):
    # R: set of rounding points enabled
    r=lambda name,v: h(v, name in R)
    def W(k):
        grp = 'w_inproj' if 'in_proj' in k else 'w_outproj' if 'out_proj' in k else 'w_lin1' if 'linear1' in k else 'w_lin2' if 'linear2' in k else 'w_fc0' if 'fc_out' in k else 'w_emb'
        if grp == 'w_inproj':
            w = sd[k]
            parts = [h(w[i*768:(i+1)*768], ('w_in_'+n) in R) for i,n in enumerate('qkv')]
            return torch.cat(parts, 0)
        return r(grp, sd[k])
    def mlp_embed(name, xx):
        hh = xx @ sd[name+'.0.weight'].t() + sd[name+'.0.bias']
        hh = F.silu(F.layer_norm(hh,(D,),sd[name+'.1.weight'],sd[name+'.1.bias'],1e-5))
        hh = r('emb_h', hh)
        return hh @ W(name+'.3.weight').t() + sd[name+'.3.bias']
    c = O.embed_mlp(sd,'time_embed',O.sincos_embedding(t)).unsqueeze(1)
    xx = mlp_embed('p_embed', x) + c
    B,L,_=xx.shape
    for i in range(12):
        p=f'net.layers.{i}'
        hn = r('xn', F.layer_norm(xx,(D,),sd[p+'.norm1.weight'],sd[p+'.norm1.bias'],1e-5))
        qkv = r('qkv', hn @ W(p+'.self_attn.in_proj_weight').t() + sd[p+'.self_attn.in_proj_bias'])
        q,k,v = qkv.split(D,-1)
        q=q.view(B,L,NH,64).transpose(1,2);k=k.view(B,L,NH,64).transpose(1,2);v=v.view(B,L,NH,64).transpose(1,2)
        s=(q@k.transpose(-1,-2))/8
        m=s.max(-1,keepdim=True).values
        pe=r('p', torch.exp(s-m))
        a=(pe@v)/torch.exp(s-m).sum(-1,keepdim=True)
        a=r('ao', a.transpose(1,2).reshape(B,L,D))
        xx = xx + a @ W(p+'.self_attn.out_proj.weight').t() + sd[p+'.self_attn.out_proj.bias']
        hn = r('xn', F.layer_norm(xx,(D,),sd[p+'.norm2.weight'],sd[p+'.norm2.bias'],1e-5))
        ff = r('hff', torch.relu(hn @ W(p+'.linear1.weight').t() + sd[p+'.linear1.bias']))
        xx = xx + ff @ W(p+'.linear2.weight').t() + sd[p+'.linear2.bias']
    xx = r('xn_final', F.layer_norm(xx,(D,),sd['net.norm.weight'],sd['net.norm.bias'],1e-5))
    hh = xx @ W('fc_out.0.weight').t() + sd['fc_out.0.bias']
    hh = r('fc_h', F.silu(F.layer_norm(hh,(D,),sd['fc_out.1.weight'],sd['fc_out.1.bias'],1e-5)))
    return hh @ sd['fc_out.3.weight'].t() + sd['fc_out.3.bias']
sd = synth_state_dict(denoiser_spec('surfpos',False), seed=7)
g=torch.Generator().manual_seed(2)
x=torch.randn(4,50,6,generator=g); t=torch.tensor([74])
rel=lambda a,b: float((a-b).double().norm()/b.double().norm())
with torch.no_grad():
    ref=run(sd,x,t,set())
    print('check vs oracle', rel(ref, O.surfpos_forward(sd,x,t)))
    for p in ['w_in_q','w_in_k','w_in_v','w_outproj']:
        print(p, '%.3e'%rel(run(sd,x,t,{p}),ref))
    base={'w_lin1','w_lin2','xn','qkv','p','ao','hff','w_emb','emb_h'}   # what precision 1 leaves rounded
    print('p1 (in,out split)', '%.3e'%rel(run(sd,x,t,base),ref))
    print('p1 but q,k single', '%.3e'%rel(run(sd,x,t,base|{'w_in_q','w_in_k'}),ref))
    print('p1 but q,k,v single (only out split)', '%.3e'%rel(run(sd,x,t,base|{'w_in_q','w_in_k','w_in_v'}),ref))
    print('p0', '%.3e'%rel(run(sd,x,t,base|{'w_in_q','w_in_k','w_in_v','w_outproj','w_fc0','xn_final','fc_h'}),ref))
