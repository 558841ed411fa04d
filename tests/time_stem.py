"""Developer tool: time the full-size stem (B=8, 800x1333) as the sliding-window GEMM + its weight gradient + the pooling kernels."""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

B, H, W = 8, 800, 1333
H2, W2 = (H + 1) // 2, (W + 1) // 2
HP, WP = H2 + 3, W2 + 3
M = B * HP * WP
BF = torch.bfloat16
img = torch.randn(B, H, W, 3, device="cuda")
s2d = torch.zeros(M + 3 * WP + 8, 16, dtype=BF, device="cuda")
w16 = (torch.randn(64, 256, device="cuda") * 0.08).to(BF)
shift = torch.zeros(64, device="cuda")
y = torch.zeros(B, HP, WP, 64, dtype=BF, device="cuda")
dy = torch.zeros(B, HP, WP, 64, dtype=BF, device="cuda")
dy[:, :H2, :W2] = torch.randn(B, H2, W2, 64, device="cuda").to(BF)
dW = torch.zeros(64, 256, device="cuda")
db = torch.zeros(64, device="cuda")
oh, ow = (H2 - 1) // 2 + 1, (W2 - 1) // 2 + 1
p1 = torch.zeros(B, oh, ow, 64, dtype=BF, device="cuda")
a1 = torch.zeros(B, oh, ow, 64, dtype=torch.uint8, device="cuda")
g = ops.plain_geom(M, 256)


def t(name, fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / n * 1e3:.1f} us")


t("s2d(pad)", lambda: ops.image_to_s2d16(img, s2d, B, H, W, 2, 2, HP, WP))
for mode in (0, 2):
    ops.set_tc_persistent(mode)
    t(f"stem sliding GEMM (persistent={mode})", lambda: ops.igemm(s2d, w16, M, 64, 256, 16, 256, g, bias=shift, relu=True, C=y, ldc=64, a_kb_rows=WP))
ops.set_tc_persistent(1)
t("stem sliding GEMM (auto)", lambda: ops.igemm(s2d, w16, M, 64, 256, 16, 256, g, bias=shift, relu=True, C=y, ldc=64, a_kb_rows=WP))
t("maxpool fwd", lambda: ops.maxpool_fwd(y, p1, a1, B, H2, W2, 64, oh, ow, XH=HP, XW=WP))
t("maxpool bwd", lambda: ops.maxpool_bwd(p1, a1, dy, B, H2, W2, 64, oh, ow, XH=HP, XW=WP))
t("stem wgrad", lambda: ops.wgrad(s2d, 16, dy, 64, M, 64, 256, g, dW, 256, dbias=db, a_kb_rows=WP, k_mask=True))
