"""Developer tool: the CTA-pair persistent GEMM (cta_group::2) against the single-CTA persistent kernel on the tensor-bound shapes of
the backbone.  python tests/time_pair.py"""
import sys

import torch

sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops  # noqa: E402

BF = torch.bfloat16


def conv_geom(B, ih, iw, cin, oh, ow, kh, kw, stride, pad, mode=0):
    return dict(batch=B, IH=ih, IW=iw, Cin=cin, OH=oh, OW=ow, KH=kh, KW=kw, stride=stride, pad=pad, mode=mode)


def timeit(fn, n=30):
    for _ in range(5):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def run(name, M, N, K, conv=None, resid=False):
    g = torch.Generator().manual_seed(0)
    if conv:
        B, H, W, C = conv
        A = torch.randn(B, H, W, C, generator=g).cuda().to(BF)
        geom = conv_geom(B, H, W, C, H, W, 3, 3, 1, 1)
        lda = C
    else:
        A = torch.randn(M, K, generator=g).cuda().to(BF)
        geom = ops.plain_geom(M, K)
        lda = K
    Wt = (torch.randn(N, K, generator=g) * K ** -0.5).cuda().to(BF)
    bias = torch.randn(N, generator=g).cuda()
    y = torch.empty(M, N, dtype=BF, device="cuda")
    kw = dict(bias=bias, relu=True)
    if resid:
        kw.update(residual=torch.randn(M, N, generator=g).cuda().to(BF), ldr=N)
    out = []
    for pair in (0, 2):
        old = ops.set_tc_pair(pair)
        t = timeit(lambda: ops.igemm(A, Wt, M, N, K, lda, K, geom, C=y, ldc=N, **kw))
        if "-v" in sys.argv:
            from torch.profiler import ProfilerActivity, profile
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                ops.igemm(A, Wt, M, N, K, lda, K, geom, C=y, ldc=N, **kw)
                torch.cuda.synchronize()
            print("   ", pair, [e.key[:60] for e in prof.key_averages()], flush=True)
        ops.set_tc_pair(old)
        out.append(t)
    fl = 2.0 * M * N * K
    print(f"{name:44s} single {out[0]:7.1f} us {fl / out[0] / 1e6:7.0f} TF/s | pair {out[1]:7.1f} us {fl / out[1] / 1e6:7.0f} TF/s  x{out[0] / out[1]:.2f}", flush=True)


if __name__ == "__main__":
    run("conv3x3 256->256 layer3 (8x50x84)", 8 * 50 * 84, 256, 2304, conv=(8, 50, 84, 256))
    run("  same shape, plain operand (no im2col)", 8 * 50 * 84, 256, 2304)
    run("  same, M = 148 x 2 x 128 (two full waves)", 148 * 2 * 128, 256, 2304)
    run("  same, M = 148 x 8 x 128 (eight full waves)", 148 * 8 * 128, 256, 2304)
    run("conv3x3 256->256, 8 x 148 x 128 pixels", 8 * 148 * 128, 256, 2304, conv=(8, 148, 128, 256))
    run("conv3x3 128->128 layer2 (8x100x167)", 8 * 100 * 167, 128, 1152, conv=(8, 100, 167, 128))
    run("conv3x3 512->512 layer4 (8x25x42)", 8 * 25 * 42, 512, 4608, conv=(8, 25, 42, 512))
    run("1x1 1024->256 layer3 conv1 (M=33600)", 33600, 256, 1024)
    run("1x1 256->1024 layer3 conv3 + residual", 33600, 1024, 256, resid=True)
    run("1x1 512->128 layer2 conv1 (M=133600)", 133600, 128, 512)
    run("1x1 2048->512 layer4 conv1 (M=8400)", 8400, 512, 2048)
    run("1x1 512->2048 layer4 conv3 + residual", 8400, 2048, 512, resid=True)
    run("1x1 1024->2048 layer4 shortcut (dgrad-like)", 8400, 1024, 2048)
    run("FFN2 2048->256 (M=8400)", 8400, 256, 2048, resid=True)
    run("plain 16384 x 4096 x 4096", 16384, 4096, 4096)
