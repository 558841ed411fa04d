"""Developer probe: how exact is the fp32 accumulation of tcgen05.mma in tensor memory?  bf16-exact inputs (so the products are
exact), fp32 output straight from the accumulator (Cf), against an fp64 product -- the error is the accumulation alone."""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

torch.manual_seed(0)
for M, N, K in ((512, 256, 256), (512, 256, 2304), (512, 512, 4608), (4096, 64, 576)):
    A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    W = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    Cf = torch.empty(M, N, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), Cf=Cf, ldcf=N, force_tc=0)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t()
    f32 = (A.float() @ W.float().t()).double()
    e = (Cf.double() - ref).abs()
    e32 = (f32 - ref).abs()
    print(f"M={M} N={N} K={K}: tcgen05 max|err|/max|ref| = {float(e.max() / ref.abs().max()):.3e}  rms rel = {float(e.norm() / ref.norm()):.3e}"
          f"   (cuBLAS fp32 on the same inputs: max {float(e32.max() / ref.abs().max()):.3e} rms {float(e32.norm() / ref.norm()):.3e})"
          f"  mean signed err / rms = {float((Cf.double() - ref).mean() / ref.pow(2).mean().sqrt()):.3e}")
