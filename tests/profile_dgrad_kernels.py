"""Developer tool: the two layer1 data-gradient launches of the bench's `rooflines` (dgrad_conv1x1_layer1: streaming tcgen05 GEMM
256 -> 64 with a 1-bit ReLU mask; dgrad_conv3x3_layer1_halo: halo conv, flipped taps, 1-bit mask) at B=8, 200x334, a few times each
with the L2 flushed in between, so that `ncu --set full -k regex:gemm_stream|halo` can capture them in isolation."""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

BF = torch.bfloat16
B, H, W, C = 8, 200, 334, 64
M = B * H * W
flush = torch.empty(1 << 28, dtype=torch.float32, device="cuda")
# (a) dX[M,64] = (dY[M,256] . W[256,64]) * relu'(bits)
dy = torch.randn(M, 256, device="cuda").to(BF)
wd = (torch.randn(64, 256, device="cuda") / 16).to(BF)
bits = torch.randint(0, 256, (M, 8), dtype=torch.uint8, device="cuda")
dx = torch.empty(M, 64, dtype=BF, device="cuda")
for _ in range(4):
    flush.fill_(1.0)
    ops.igemm(dy, wd, M, 64, 256, 256, 256, ops.plain_geom(M, 256), mask_bits=bits, ldmb=8, mask_scale=1.0, C=dx, ldc=64)
# (b) 3x3 64 -> 64 data gradient on the halo kernel
dy3 = torch.randn(B, H, W, C, device="cuda").to(BF)
w3 = (torch.randn(C, 9 * C, device="cuda") / 24).to(BF)
g = dict(batch=B, IH=H, IW=W, Cin=C, OH=H, OW=W, KH=3, KW=3, stride=1, pad=1, mode=1)
dx3 = torch.empty(M, C, dtype=BF, device="cuda")
for _ in range(4):
    flush.fill_(1.0)
    ops.igemm(dy3, w3, M, C, 9 * C, C, 9 * C, g, mask_bits=bits, ldmb=8, mask_scale=1.0, C=dx3, ldc=C)
torch.cuda.synchronize()
print("done", float(dx[:1000].float().abs().mean()), float(dx3[:1000].float().abs().mean()))
