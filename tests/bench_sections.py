"""Developer tool (not a test): per-section timing of the train step and a GEMM shape sweep, igemm (mma.sync) vs
gemm_tc (tcgen05).  python tests/bench_sections.py [B H W]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import detr_tensorflow_b200 as D  # noqa: E402
from detr_tensorflow_b200 import ops  # noqa: E402

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 800, 1333)


def time_fn(fn, iters=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3      # us


def gemm_sweep():
    shapes = [(8400, 256, 256), (8400, 512, 256), (8400, 2048, 256), (8400, 256, 2048), (800, 256, 256), (800, 2048, 256),
              (800, 256, 2048), (534400, 64, 64), (534400, 256, 64), (534400, 64, 256), (133600, 128, 512), (133600, 512, 128),
              (33600, 256, 1024), (33600, 1024, 256), (8400, 512, 2048), (8400, 2048, 512), (8400, 256, 2048)]
    out = []
    for (M, N, K) in shapes:
        A = torch.randn(M, K, device="cuda").to(torch.bfloat16)
        Wt = torch.randn(N, K, device="cuda").to(torch.bfloat16)
        C = torch.empty(M, N, dtype=torch.bfloat16, device="cuda")
        R = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(N, device="cuda")
        row = {"M": M, "N": N, "K": K}
        for name, tc in (("tc_1tile", 0), ("tc_persist", 0)):
            ops.set_tc_persistent(1 if name == "tc_persist" else 0)
            try:
                us = time_fn(lambda: ops.igemm(A, Wt, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, residual=R, ldr=N, C=C, ldc=N, force_tc=tc))
                row[name + "_us"] = round(us, 1)
                row[name + "_tflops"] = round(2.0 * M * N * K / us / 1e6, 1)
                row[name + "_gbs"] = round((M * K + N * K + 2 * M * N) * 2 / us / 1e3, 0)
            except Exception as e:      # noqa
                row[name + "_err"] = str(e)[:100]
        out.append(row)
        print(json.dumps(row), flush=True)
    return out


def tcp_sweep():
    """one-tile kernel vs the persistent kernel (forced tile widths) on the step's big GEMM / conv shapes"""
    import torch.nn.functional as F  # noqa
    gemms = [(534400, 256, 64, "r"), (534400, 256, 64, "rm"), (534400, 64, 256, ""), (534400, 64, 64, ""), (133600, 512, 128, "r"),
             (133600, 512, 128, "rm"), (133600, 128, 512, ""), (33600, 1024, 256, "r"), (33600, 256, 1024, ""), (8400, 2048, 512, "r"),
             (8400, 512, 2048, ""), (8400, 2048, 256, ""), (8400, 256, 2048, "r"), (8400, 256, 256, "r")]
    convs = [(8, 200, 334, 64), (8, 100, 167, 128), (8, 50, 84, 256), (8, 25, 42, 512)]
    def run(label, fn, flops, byts):
        row = {"shape": label}
        for name, mode, bn in (("1tile", 0, 0), ("p64", 2, 64), ("p128", 2, 128), ("p256", 2, 256), ("auto", 1, 0)):
            ops.set_tc_persistent(mode)
            try:
                us = time_fn(lambda: fn(bn), iters=10)
                row[name] = f"{us:.1f}us {flops / us / 1e6:.0f}TF {byts / us / 1e3:.0f}GB/s"
            except Exception as e:      # noqa
                row[name] = "n/a"
        ops.set_tc_persistent(0)
        print(json.dumps(row), flush=True)
    for (M, N, K, epi) in gemms:
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
        byts = (M * K + N * K + (1 + len(epi)) * M * N) * 2
        run(f"gemm {M}x{N}x{K} {epi}", lambda bn: ops.igemm(A, Wt, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, C=C, ldc=N, force_tc=bn, **kw),
            2.0 * M * N * K, byts)
    for (Bn, Hh, Ww, Cc) in convs:
        M, K = Bn * Hh * Ww, 9 * Cc
        x = torch.randn(M, Cc, device="cuda").to(torch.bfloat16)
        Wt = torch.randn(Cc, K, device="cuda").to(torch.bfloat16)
        C = torch.empty(M, Cc, dtype=torch.bfloat16, device="cuda")
        Mk = torch.randn(M, Cc, device="cuda").to(torch.bfloat16)
        bias = torch.zeros(Cc, device="cuda")
        g = dict(batch=Bn, IH=Hh, IW=Ww, Cin=Cc, OH=Hh, OW=Ww, KH=3, KW=3, stride=1, pad=1, mode=0)
        run(f"conv3x3 fwd {Cc}ch {Hh}x{Ww}", lambda bn: ops.igemm(x, Wt, M, Cc, K, Cc, K, g, bias=bias, relu=True, C=C, ldc=Cc, force_tc=bn),
            2.0 * M * Cc * K, (2 * M * Cc + Cc * K) * 2)
        g2 = dict(g, mode=1)
        run(f"conv3x3 dgrad {Cc}ch {Hh}x{Ww}", lambda bn: ops.igemm(x, Wt, M, Cc, K, Cc, K, g2, mask=Mk, ldm=Cc, mask_scale=1.0, C=C, ldc=Cc, force_tc=bn),
            2.0 * M * Cc * K, (3 * M * Cc + Cc * K) * 2)


def sections(tc):
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, seed=0)
    eng = model.engine
    g = torch.Generator().manual_seed(0)
    img = torch.randn(B, H, W, 3, generator=g)
    tb = torch.zeros(B, 100, 4)
    tc_ = torch.zeros(B, 100, 1, dtype=torch.int64)
    for b in range(B):
        tb[b, 0, 0] = 20
        tb[b, 1:21, :2] = torch.rand(20, 2, generator=g) * 0.8 + 0.1
        tb[b, 1:21, 2:] = torch.rand(20, 2, generator=g) * 0.48 + 0.02
        tc_[b, 1:21, 0] = torch.randint(0, 91, (20,), generator=g)
    old = ops.set_tc(1 if tc else 0)
    oldw = ops.set_tc_wgrad(1 if tc >= 2 else 0)
    oldp = ops.set_tc_persistent(1 if tc >= 3 else 0)      # 3: persistent kernel, auto policy
    try:
        eng.forward(img, training=True)
        eng.set_targets(tb, tc_)
        eng.set_lrs(1e-5, 1e-4)
        eng.set_enabled(True, True)
        eng.train_step(91, 0.1)
        res = eng.profile_step(91, 0.1)
        res["total"] = sum(res.values())
        res["loss"] = float(eng.a["total"][0])
    finally:
        ops.set_tc(old)
        ops.set_tc_wgrad(oldw)
        ops.set_tc_persistent(oldp)
    print(json.dumps({"tc": tc, "sections_ms": {k: round(v, 3) for k, v in res.items()}}), flush=True)
    del model, eng
    torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[4] if len(sys.argv) > 4 else "all"
    if what in ("all", "gemm"):
        gemm_sweep()
    if what == "tcp":
        tcp_sweep()
    if what in ("all", "sections"):
        for mode in (2, 3):
            try:
                sections(mode)
            except Exception as e:      # noqa
                print(json.dumps({"tc": mode, "error": str(e)[:300]}), flush=True)
