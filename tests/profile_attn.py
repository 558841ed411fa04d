"""Developer tool (not a test): the encoder's self-attention shape (B=8, H=8, S=1050, dh=32, dropout 0.1) through
attn_fwd / attn_bwd a few times, for `ncu --set full -k regex:attn_`."""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

B, H, S, d = 8, 8, int(sys.argv[1]) if len(sys.argv) > 1 else 1050, 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
scale = 32 ** -0.5
g = torch.Generator(device="cuda").manual_seed(0)
qk = torch.randn(B * S, 2 * d, device="cuda", generator=g).to(torch.bfloat16)
v = torch.randn(B * S, d, device="cuda", generator=g).to(torch.bfloat16)
o = torch.zeros(B * S, d, dtype=torch.bfloat16, device="cuda")
do = torch.randn(B * S, d, device="cuda", generator=g).to(torch.bfloat16)
lse = torch.zeros(B * H * S, dtype=torch.float32, device="cuda")
delta = torch.zeros(B * H * S, dtype=torch.float32, device="cuda")
dqk = torch.zeros(B * S, 2 * d, dtype=torch.bfloat16, device="cuda")
dv = torch.zeros(B * S, d, dtype=torch.bfloat16, device="cuda")
seed_dev = torch.tensor([7], dtype=torch.int64, device="cuda")
kw = dict(drop_p=0.1, seed=3, site=11, seed_ptr=seed_dev)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
for i in range(reps):
    ev[0].record()
    ops.attn_fwd(qk, qk[:, d:], v, 2 * d, 2 * d, d, o, d, lse, B, H, S, S, scale, **kw)
    ev[1].record()
    ops.attn_bwd(qk, qk[:, d:], v, o, do, 2 * d, 2 * d, d, d, d, lse, delta, dqk, dqk[:, d:], dv, 2 * d, 2 * d, d, B, H, S, S, scale, **kw)
    ev[2].record()
torch.cuda.synchronize()
print("fwd us", ev[0].elapsed_time(ev[1]) * 1e3, "bwd us", ev[1].elapsed_time(ev[2]) * 1e3)
