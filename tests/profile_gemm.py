"""Developer tool (not a test): run one GEMM shape through the tcgen05 kernels a few times (for ncu).
python tests/profile_gemm.py M N K epi mode bn     epi: "", "r", "rm"; mode: 0 one-tile, 2 persistent; bn: 0/64/128/256"""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
epi, mode, bn = sys.argv[4].strip("-"), int(sys.argv[5]), int(sys.argv[6])
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
Wt = torch.randn(N, K, device="cuda").to(torch.bfloat16)
C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
R = torch.randn(M, N, device="cuda").to(torch.bfloat16)
Mk = torch.randn(M, N, device="cuda").to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
kw = {}
if "r" in epi:
    kw.update(residual=R, ldr=N)
if "m" in epi:
    kw.update(mask=Mk, ldm=N, mask_scale=1.0)
ops.set_tc_persistent(mode)
for _ in range(3):
    ops.igemm(A, Wt, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, C=C, ldc=N, force_tc=bn, **kw)
torch.cuda.synchronize()
print("ok")
