"""Developer tool: the weight-gradient kernels on layer1's shapes (<= 64 output channels).  python tests/time_wgrad.py"""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

BF = torch.bfloat16


def run(name, B, H, W, cin, cout, k, pad):
    M, K = B * H * W, k * k * cin
    g = torch.Generator().manual_seed(0)
    x = torch.randn(B, H, W, cin, generator=g).cuda().to(BF)
    dy = torch.randn(M, cout, generator=g).cuda().to(BF)
    geom = dict(batch=B, IH=H, IW=W, Cin=cin, OH=H, OW=W, KH=k, KW=k, stride=1, pad=pad, mode=0)
    dW = torch.zeros(cout, K, dtype=torch.float32, device="cuda")
    db = torch.zeros(cout, dtype=torch.float32, device="cuda")
    fn = lambda: ops.wgrad(x, cin, dy, cout, M, cout, K, geom, dW, K, dbias=db, force_tc=True)
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        fn()
        fn()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 250)
    ts.sort()
    byts = (x.numel() + dy.numel()) * 2
    print(f"{name:34s} {ts[10]:7.1f} us  {byts / ts[10] / 1e3:6.0f} GB/s  {2.0 * M * cout * K / ts[10] / 1e6:6.0f} TF/s", flush=True)


run("layer1 3x3 64->64 (8x200x334)", 8, 200, 334, 64, 64, 3, 1)
run("layer1 1x1 256->64", 8, 200, 334, 256, 64, 1, 0)
run("layer1 1x1 64->64 (block0 conv1)", 8, 200, 334, 64, 64, 1, 0)
run("layer1 1x1 64->256 (general kernel)", 8, 200, 334, 64, 256, 1, 0)
