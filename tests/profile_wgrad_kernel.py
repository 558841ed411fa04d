"""Developer tool: launch the tcgen05 weight-gradient kernel on the bench's `wgrad_conv3x3` shape (3x3 256->256 conv of layer3 at
B=8, 50x84) and on the layer1 1x1 shape a few times so that `ncu --set full -k regex:wgrad_tc` can capture them in isolation."""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "3x3"
if which == "3x3":
    B, H, W, C, N, k, pad = 8, 50, 84, 256, 256, 3, 1
else:                                                      # layer1 conv3: 1x1 64 -> 256 at 200x334
    B, H, W, C, N, k, pad = 8, 200, 334, 64, 256, 1, 0
x = torch.randn(B, H, W, C, device="cuda").to(torch.bfloat16)
dy = torch.randn(B, H, W, N, device="cuda").to(torch.bfloat16)
dW = torch.zeros(N, k * k * C, device="cuda")
db = torch.zeros(N, device="cuda")
g = dict(batch=B, IH=H, IW=W, Cin=C, OH=H, OW=W, KH=k, KW=k, stride=1, pad=pad, mode=0)
for _ in range(6):
    ops.wgrad(x, C, dy, N, B * H * W, N, k * k * C, g, dW, k * k * C, dbias=db)
torch.cuda.synchronize()
print("done", float(dW.abs().mean()))
