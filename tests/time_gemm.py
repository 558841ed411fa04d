"""Developer tool: time one GEMM shape through detrb_igemm (auto policy; env switches select kernel variants).
python tests/time_gemm.py M N K epi [mode bn nobias]  (epi: -, r, m, rm; mode 0 one-tile / 2 persistent; bn 0|64|128|256)"""
import os
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

M, N, K = (int(x) for x in sys.argv[1:4])
epi = sys.argv[4].strip("-")
A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
Wt = torch.randn(N, K, device="cuda").to(torch.bfloat16)
C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
R = torch.randn(M, N, device="cuda").to(torch.bfloat16)
Mk = torch.randn(M, N, device="cuda").to(torch.bfloat16)
bias = torch.zeros(N, device="cuda")
mode = int(sys.argv[5]) if len(sys.argv) > 5 else -1
bn = int(sys.argv[6]) if len(sys.argv) > 6 else 0
if len(sys.argv) > 7 and sys.argv[7] == "nobias":
    bias = None
import os
if os.environ.get("TMAEPI") == "0":
    from detr_tensorflow_b200 import _lib
    _lib.lib().detrb_set_tc_tma_epilogue(0)
kw = {}
if mode >= 0:
    ops.set_tc_persistent(mode)
    kw["force_tc"] = bn
if "r" in epi:
    kw.update(residual=R, ldr=N)
if "m" in epi:
    kw.update(mask=Mk, ldm=N, mask_scale=1.0)
if "b" in epi:                      # 1-bit mask instead of the bf16 one
    Mb = torch.randint(0, 256, (M, N // 8), dtype=torch.uint8, device="cuda")
    kw.update(mask_bits=Mb, ldmb=N // 8, mask_scale=1.0)
if "o" in epi:                      # also write the 1-bit ReLU mask of the result
    Ob = torch.empty(M, N // 8, dtype=torch.uint8, device="cuda")
    kw.update(out_bits=Ob, ldob=N // 8)
if os.environ.get("STREAM") is not None:
    ops.set_tc_stream(int(os.environ["STREAM"]))
fn = lambda: ops.igemm(A, Wt, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, C=C, ldc=N, **kw)
for _ in range(3):
    fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize()
e0.record()
for _ in range(20):
    fn()
e1.record()
torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
byts = (M * K + N * K + (1 + len(epi.replace("b", "").replace("o", ""))) * M * N) * 2 + (("b" in epi) + ("o" in epi)) * M * N // 8
ref = torch.relu(A[:256].float() @ Wt.float().t() + (R[:256].float() if "r" in epi else 0))
if "m" in epi:
    ref = torch.where(Mk[:256].float() > 0, ref, torch.zeros_like(ref))
if "b" in epi:
    bit = ((Mb[:256, :, None] >> torch.arange(8, device="cuda", dtype=torch.uint8)[None, None, :]) & 1).reshape(256, N) > 0
    ref = torch.where(bit, ref, torch.zeros_like(ref))
err = float((C[:256].float() - ref).abs().max() / (ref.abs().max() + 1e-9))
print(f"{M}x{N}x{K} {epi or '-'}: {us:.1f} us  {byts / us / 1e3:.0f} GB/s  {2.0 * M * N * K / us / 1e6:.0f} TF/s  relerr {err:.4f}")
if os.environ.get("DETRB_SO"):
    import ctypes
    from detr_tensorflow_b200 import _lib as L_
    buf = (ctypes.c_ulonglong * 16)()
    L_.lib().detrb_trace_read(buf, 1)
    fn()
    L_.lib().detrb_trace_read(buf, 1)
    n = max(1, buf[0])
    names = ["ctas", "entry->sync(alloc,init)", "pdl wait", "pdl->mma issued", "pdl->acc complete", "acc->inputs landed",
             "ld+math+sts", "fence+bar", "store+drain", "cta life"]
    print("   trace (cycles/CTA): " + ", ".join(f"{names[i]}={buf[i] / n:.0f}" for i in range(1, 10)) + f", ctas={buf[0]}")
