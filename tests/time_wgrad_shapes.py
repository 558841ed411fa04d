"""Developer tool: every weight-gradient shape of the ResNet-50 backbone at B=8, 800x1333, timed alone (CUDA events, three
launches per sample, median of 15) against its own bounds: tensor (2 M N K / 1654.8 TF/s) and HBM (A + dY once / 6552.6 GB/s).
python tests/time_wgrad_shapes.py"""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

BF = torch.bfloat16


def run(name, count, B, IH, IW, cin, cout, k, stride):
    pad = k // 2
    OH, OW = (IH + 2 * pad - k) // stride + 1, (IW + 2 * pad - k) // stride + 1
    M, K = B * OH * OW, k * k * cin
    x = torch.randn(B, IH, IW, cin, device="cuda").to(BF)
    dy = torch.randn(M, cout, device="cuda").to(BF)
    geom = dict(batch=B, IH=IH, IW=IW, Cin=cin, OH=OH, OW=OW, KH=k, KW=k, stride=stride, pad=pad, mode=0)
    dW = torch.zeros(cout, K, dtype=torch.float32, device="cuda")
    db = torch.zeros(cout, dtype=torch.float32, device="cuda")
    fn = lambda: ops.wgrad(x, cin, dy, cout, M, cout, K, geom, dW, K, dbias=db)
    flops, byts = 2.0 * M * cout * K, (x.numel() + dy.numel()) * 2.0
    bound = max(flops / 1654.8e6, byts / 6552.6e3)
    res = []
    for mode in MODES:
        ops.set_wgrad_tile(mode)
        for _ in range(3):
            fn()
        ts = []
        for _ in range(15):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            fn()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1000 / 3)
        ts.sort()
        res.append(ts[7])
    ops.set_wgrad_tile(0)
    print(f"{name:36s} x{count:2d} bound {bound:6.1f} us | " + "  ".join(f"tile{m}: {t:6.1f}" for m, t in zip(MODES, res)) +
          f" | best {flops / min(res) / 1e6:6.0f} TF/s {byts / min(res) / 1e3:6.0f} GB/s", flush=True)
    return count * res[0], count * min(res)


MODES = [int(a) for a in sys.argv[1:]] or [0, 1, 2, 3, 4]      # 0 policy, 1 128x128, 2 128x256, 3 256x128, 4 256x256
B = 8
tot = [0.0, 0.0]
for args in [
    ("layer1 1x1 64->64", 1, B, 200, 334, 64, 64, 1, 1), ("layer1 3x3 64->64", 3, B, 200, 334, 64, 64, 3, 1),
    ("layer1 1x1 64->256", 4, B, 200, 334, 64, 256, 1, 1), ("layer1 1x1 256->64", 2, B, 200, 334, 256, 64, 1, 1),
    ("layer2 1x1 256->128 @200", 1, B, 200, 334, 256, 128, 1, 1), ("layer2 3x3/2 128->128", 1, B, 200, 334, 128, 128, 3, 2),
    ("layer2 1x1/2 256->512", 1, B, 200, 334, 256, 512, 1, 2), ("layer2 1x1 128->512", 4, B, 100, 167, 128, 512, 1, 1),
    ("layer2 1x1 512->128", 3, B, 100, 167, 512, 128, 1, 1), ("layer2 3x3 128->128", 3, B, 100, 167, 128, 128, 3, 1),
    ("layer3 1x1 512->256 @100", 1, B, 100, 167, 512, 256, 1, 1), ("layer3 3x3/2 256->256", 1, B, 100, 167, 256, 256, 3, 2),
    ("layer3 1x1/2 512->1024", 1, B, 100, 167, 512, 1024, 1, 2), ("layer3 1x1 256->1024", 6, B, 50, 84, 256, 1024, 1, 1),
    ("layer3 1x1 1024->256", 5, B, 50, 84, 1024, 256, 1, 1), ("layer3 3x3 256->256", 5, B, 50, 84, 256, 256, 3, 1),
    ("layer4 1x1 1024->512 @50", 1, B, 50, 84, 1024, 512, 1, 1), ("layer4 3x3/2 512->512", 1, B, 50, 84, 512, 512, 3, 2),
    ("layer4 1x1/2 1024->2048", 1, B, 50, 84, 1024, 2048, 1, 2), ("layer4 1x1 512->2048", 3, B, 25, 42, 512, 2048, 1, 1),
    ("layer4 1x1 2048->512", 2, B, 25, 42, 2048, 512, 1, 1), ("layer4 3x3 512->512", 2, B, 25, 42, 512, 512, 3, 1),
    ("input_proj 2048->256", 1, B, 25, 42, 2048, 256, 1, 1), ("encoder linear 256->256 (M=8400)", 24, 1, 1, 8400, 256, 256, 1, 1),
    ("encoder ffn1 256->2048", 6, 1, 1, 8400, 256, 2048, 1, 1), ("encoder ffn2 2048->256", 6, 1, 1, 8400, 2048, 256, 1, 1),
]:
    a, b = run(*args)
    tot[0] += a
    tot[1] += b
print(f"sum with the first mode {tot[0]:.0f} us, with the best tile per shape {tot[1]:.0f} us")
