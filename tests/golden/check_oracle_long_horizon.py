#!/usr/bin/env python
"""Oracle vs the REAL reference over the long horizon, on the CPU: the oracle's loop restatement (oracle/ref_loop_torch.py) runs the
2 000-iteration block reconstruction of ResNet-18 layer1.0 that tests/golden/long_horizon.npz recorded from the reference
(make_golden_round2.py), with the block inputs regenerated here from the oracle's quantised prefix and the index stream from the seed.
Result (this container, 8 cores, 316 s): hard codes of conv1 / conv2 100 % identical, alphas identical after rounding to fp16 — the
restatement follows the reference's trajectory exactly, so every code the GPU run differs in (profiles/r02_parity_report.json) comes
from cuDNN-vs-CPU convolution arithmetic, not from the oracle or the loop logic. Too slow for the default CPU suite; run by hand:

    python tests/golden/check_oracle_long_horizon.py
"""
import sys, time, numpy as np, torch
import os
ROOT=os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0,ROOT); sys.path.insert(0,os.path.join(ROOT,'tests'))
from oracle import ref_loop_torch as R
from oracle import ssq_oracle as O
from shiftedscalequantization_b200 import quant as Q, zoo
from shiftedscalequantization_b200.engine import index_table
g=np.load(os.path.join(ROOT,'tests','golden','long_horizon.npz'))
AQ={'n_bits':4,'channel_wise':False,'scale_method':'mse','leaf_param':True}
torch.manual_seed(1005)
cnn=zoo.resnet18(num_classes=10).eval()
qnn=Q.QuantModel(cnn,{'n_bits':2,'channel_wise':True,'scale_method':'max'},dict(AQ)).eval()
qnn.set_first_last_layer_to_8bit()
cali=torch.randn(64,3,32,32)
assert np.array_equal(cali.reshape(-1)[:64].numpy(), g['cali_probe'])
block=qnn.model.layer1[0]
mods=[m for m in qnn.modules() if isinstance(m,Q.QuantModule)]
def capture(unit):
    seen={}
    h=unit.register_forward_hook(lambda m,i,o: seen.update(inp=i[0].detach().clone(), out=o.detach().clone()))
    qnn.set_quant_state(False,False)
    with torch.no_grad(): qnn(cali)
    h.remove(); return seen['inp'], seen['out']
def qparams(m):
    nb=m.weight_quantizer.n_bits; w=m.weight.detach().numpy(); rows=w.reshape(w.shape[0],-1)
    d,z,_=zip(*[O.max_init(r,nb) for r in rows]); shape=(-1,)+(1,)*(w.ndim-1)
    return np.array(d,np.float32).reshape(shape), np.array(z,np.float32).reshape(shape), nb
_,outs=capture(block)
saved=[m.org_weight for m in mods]; qp={}
for m in mods:
    d,z,nb=qparams(m); qp[m]=(d,z,nb)
    y,_=O.uaq_forward(m.weight.detach().numpy(),d,z,0,2**nb-1); m.org_weight=torch.from_numpy(y)
inps,_=capture(block)
named=[(n,m) for n,m in block.named_modules() if isinstance(m,Q.QuantModule)]
layers={}
for (n,m),w in zip(named,[saved[mods.index(m)] for n,m in named]):
    act={"ReLU":"relu"}.get(type(m.activation_function).__name__)
    d,z,nb=qp[m]
    layers[n]=dict(weight=w.detach(),bias=None if m.org_bias is None else m.org_bias.detach(),conv=dict(m.fwd_kwargs),act=act,delta=torch.from_numpy(d),zero_point=torch.from_numpy(z),n_levels=2**nb)
unit={"kind":"basic","layers":layers,"tail_act":"relu"}
iters=2000
torch.manual_seed(377); tab=index_table(64,32,iters)
t0=time.time()
alphas,losses=R.recon_weight_loop(unit,inps,outs,tab,iters,weight=0.01,b_range=(20,2),warmup=0.2)
print('block loop',time.time()-t0,'s')
for n in ('conv1','conv2'):
    L=layers[n]; a=alphas[n].detach()
    codes=torch.clamp(torch.floor(L['weight']/L['delta'])+(a>=0).float()+L['zero_point'],0,L['n_levels']-1).to(torch.uint8).numpy()
    ref=g[f'i{iters}.block.{n}.codes']
    agree=(codes==ref).mean(); same16=(a.numpy().astype(np.float16)==g[f'i{iters}.block.{n}.alpha16']).mean()
    print(n,'code agreement',agree,'alpha16 equal',same16)
    assert agree==1.0 and same16==1.0
