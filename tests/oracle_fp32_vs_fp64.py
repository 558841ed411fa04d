"""Developer tool (CPU): how reproducible are the train-step gradients in fp32 at all?  The oracle run in fp32 against the same oracle
in fp64 (same assignment): the forward pass agrees to ~2e-6, the gradients of the early backbone only to ~3e-3 -- a ReLU whose input
lies within the forward error of zero flips its mask, and a fraction f of flipped elements costs sqrt(f) in relative L2 norm.  This is
the floor any fp32-class implementation (the reference in TF included) sits on; results: profiles/r02_parity_errors.txt."""
import sys, torch
sys.path.insert(0,'.')
from oracle import detr_oracle as O
torch.set_num_threads(8)
def rel(a,b): return float((a.double()-b.double()).norm()/(b.double().norm()+1e-30))
P = O.init_params(seed=1)
img = torch.randn(2,160,224,3, generator=torch.Generator().manual_seed(1))
tb, tc = O.synthetic_targets(2, n=6, seed=1)
out32, t32, _, g32 = O.train_step(P, img, tb, tc)
# the fp32 run's assignment, reused in fp64
with torch.no_grad():
    pass
P64 = {k: v.double() for k,v in P.items()}
# match from fp32 outputs
import numpy as np
L=6
match = -torch.ones(L,2,100,dtype=torch.int32)
for l in range(L):
    o_ = out32 if l==L-1 else out32["aux"][l]
    for b in range(2):
        ti, pi, *_ = O.hungarian_matching(tb[b], tc[b], o_["pred_boxes"][b].detach(), o_["pred_logits"][b].detach())
        match[l,b,pi] = ti.int()
out64, t64, _, g64 = O.train_step(P64, img.double(), tb.double(), tc, match_override=match)
_, t32b, _, g32b = O.train_step(P, img, tb, tc, match_override=match)
print("forward logits fp32 vs fp64 rel:", rel(out32["pred_logits"], out64["pred_logits"]))
r = sorted((rel(g32b[n], g64[n]), n) for n in g64 if g64[n] is not None and float(g64[n].norm())>1e-9)
print("oracle fp32 vs fp64 gradients: median", r[len(r)//2], "worst", r[-3:])
bb = [x for x in r if x[1].startswith('backbone/')]
tr = [x for x in r if not x[1].startswith('backbone/')]
print("backbone median", bb[len(bb)//2], "transformer median", tr[len(tr)//2], "transformer worst", tr[-1])
