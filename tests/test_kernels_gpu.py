"""GPU parity tests, kernel by kernel, through the C ABI (libdetrb.so via detr_tensorflow_b200.ops).
Each CUDA kernel is compared with a plain PyTorch fp32 reference of the same op fed the same (bf16-rounded)
inputs; tolerances are stated per test.  Integer outputs (matcher indices) must be bit-exact."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module")
def ops():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import _lib, ops as o
    _lib.check(_lib.lib().detrb_check_device())
    return o


def dev(t):
    return t.cuda()


def rnd(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale)


def close(name, got, ref, rtol, atol):
    got, ref = got.float().cpu(), ref.float().cpu()
    err = (got - ref).abs()
    tol = atol + rtol * ref.abs()
    bad = int((err > tol).sum())
    rel = float((got - ref).norm() / (ref.norm() + 1e-20))
    assert bad == 0, f"{name}: {bad}/{err.numel()} elements out of tol; max err {float(err.max()):.4g}, rel-norm {rel:.4g}, " \
                     f"ref max {float(ref.abs().max()):.4g}, first bad idx {torch.nonzero(err > tol)[0].tolist()}"
    return rel


# ------------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,N,K", [(128, 128, 32), (300, 256, 256), (1050, 512, 256), (77, 64, 2048), (600, 92, 256),
                                   (600, 4, 256), (257, 2048, 256), (130, 256, 96)])
def test_plain_gemm(ops, M, N, K):
    A = dev(rnd(M, K, seed=1).to(BF))
    W = dev(rnd(N, K, scale=K ** -0.5, seed=2).to(BF))
    bias = dev(rnd(N, seed=3))
    C = torch.zeros(M, N, dtype=BF, device="cuda")
    Cf = torch.zeros(M, N, dtype=F32, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, C=C, ldc=N, Cf=Cf, ldcf=N)
    ref = A.float() @ W.float().t() + bias
    close("gemm f32", Cf, ref, 1e-3, 1e-3)
    close("gemm bf16", C, ref, 1e-2, 1e-2)


def test_gemm_epilogues(ops):
    M, N, K = 333, 256, 128
    A = dev(rnd(M, K, seed=1).to(BF))
    W = dev(rnd(N, K, scale=K ** -0.5, seed=2).to(BF))
    bias = dev(rnd(N, seed=3))
    res = dev(rnd(M, N, seed=4).to(BF))
    mask = dev(rnd(M, N, seed=5).to(BF))
    base = A.float() @ W.float().t() + bias
    C = torch.zeros(M, N, dtype=BF, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, relu=True, C=C, ldc=N)
    close("bias+res+relu", C, F.relu(base + res.float()), 1e-2, 1e-2)
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, mask=mask, ldm=N, mask_scale=1.25, C=C, ldc=N)
    close("mask", C, torch.where(mask.float() > 0, base * 1.25, torch.zeros_like(base)), 1e-2, 1e-2)
    Cf = torch.zeros(M, N, dtype=F32, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, sigmoid=True, Cf=Cf, ldcf=N)
    close("sigmoid", Cf, torch.sigmoid(base), 1e-3, 1e-3)
    # accumulate
    C0 = dev(rnd(M, N, seed=6).to(BF))
    C = C0.clone()
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), C=C, ldc=N, accumulate=True)
    close("accumulate", C, A.float() @ W.float().t() + C0.float(), 1e-2, 2e-2)
    # strided views: A with lda > K, W sub-block
    A2 = dev(rnd(M, 2 * K, seed=7).to(BF))
    W2 = dev(rnd(3 * N, K, scale=K ** -0.5, seed=8).to(BF))
    ops.igemm(A2[:, K:], W2[N:], M, N, K, 2 * K, K, ops.plain_geom(M, K), C=C, ldc=N)
    close("views", C, A2[:, K:].float() @ W2[N:2 * N].float().t(), 1e-2, 1e-2)


def test_gemm_dropout_matches_mask_kernel(ops):
    M, N, K = 200, 256, 64
    A = dev(rnd(M, K, seed=1).to(BF))
    W = dev(rnd(N, K, scale=K ** -0.5, seed=2).to(BF))
    res = dev(rnd(M, N, seed=4).to(BF))
    seed_dev = torch.tensor([12345], dtype=torch.int64, device="cuda")
    C = torch.zeros(M, N, dtype=BF, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), residual=res, ldr=N, drop_p=0.1, seed=77, site=5, seed_ptr=seed_dev,
              C=C, ldc=N)
    keep = torch.zeros(M, N, dtype=torch.uint8, device="cuda")
    ops.dropout_mask(keep, M, N, 0.1, 77, 5, seed_dev)
    frac = float(keep.float().mean())
    assert abs(frac - 0.9) < 0.01, frac
    ref = (A.float() @ W.float().t()) * keep.float() / 0.9 + res.float()
    close("dropout epilogue", C, ref, 1e-2, 2e-2)
    # a different device seed word gives a different mask
    seed_dev += 1
    keep2 = torch.zeros_like(keep)
    ops.dropout_mask(keep2, M, N, 0.1, 77, 5, seed_dev)
    assert float((keep2 != keep).float().mean()) > 0.1


# ------------------------------------------------------------------------------------------------ conv
def conv_geom(B, ih, iw, cin, oh, ow, kh, kw, stride, pad, mode=0):
    return dict(batch=B, IH=ih, IW=iw, Cin=cin, OH=oh, OW=ow, KH=kh, KW=kw, stride=stride, pad=pad, mode=mode)


CONVS = [  # cin, cout, k, stride, pad, H, W
    (64, 64, 1, 1, 0, 13, 19), (64, 64, 3, 1, 1, 13, 19), (128, 128, 3, 2, 1, 13, 19), (256, 512, 1, 2, 0, 13, 19),
    (64, 256, 1, 1, 0, 9, 11), (128, 64, 3, 2, 1, 12, 18),
]


def _out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W", CONVS)
def test_conv_fwd_dgrad_wgrad(ops, cin, cout, k, stride, pad, H, W):
    B = 2
    oh, ow = _out(H, k, stride, pad), _out(W, k, stride, pad)
    x = dev(rnd(B, H, W, cin, seed=1).to(BF))
    w = dev(rnd(cout, k, k, cin, scale=(k * k * cin) ** -0.5, seed=2).to(BF))      # [Cout][kh][kw][Cin]
    shift = dev(rnd(cout, seed=3))
    y = torch.zeros(B, oh, ow, cout, dtype=BF, device="cuda")
    M = B * oh * ow
    K = k * k * cin
    ops.igemm(x, w, M, cout, K, cin, K, conv_geom(B, H, W, cin, oh, ow, k, k, stride, pad), bias=shift, relu=True, C=y, ldc=cout)
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wt = w.float().permute(0, 3, 1, 2).requires_grad_(True)            # OIHW
    ref = F.conv2d(xt, wt, bias=shift, stride=stride, padding=pad)
    close("conv fwd", y, F.relu(ref).permute(0, 2, 3, 1), 1e-2, 2e-2)
    # ---- gradients
    dy = dev(rnd(B, oh, ow, cout, seed=5).to(BF))
    gx, gw = torch.autograd.grad(ref, [xt, wt], dy.float().permute(0, 3, 1, 2))
    # data gradient: weights in [Cin][tap][Cout] layout, transposed gather
    wd = w.reshape(cout, k * k, cin).permute(2, 1, 0).contiguous()
    dx = torch.zeros(B, H, W, cin, dtype=BF, device="cuda")
    Mi = B * H * W
    if k == 1 and stride > 1:
        # 1x1 strided conv: GEMM at the output resolution + strided scatter-accumulate
        base = dev(rnd(B, H, W, cin, seed=9).to(BF))
        dx = base.clone()
        ops.igemm(dy, wd, M, cin, cout, cout, cout, conv_geom(B, oh, ow, cout, oh, ow, 1, 1, 1, 0), C=dx, ldc=cin,
                  out_stride=stride, SH=H, SW=W, accumulate=True)
        close("1x1 strided dgrad scatter", dx, gx.permute(0, 2, 3, 1) + base.float(), 1e-2, 2e-2)
    else:
        ops.igemm(dy, wd, Mi, cin, k * k * cout, cout, k * k * cout, conv_geom(B, oh, ow, cout, H, W, k, k, stride, pad, mode=1),
                  C=dx, ldc=cin)
        close("conv dgrad", dx, gx.permute(0, 2, 3, 1), 1e-2, 2e-2)
    # weight gradient
    scale = dev(rnd(cout, seed=6).abs() + 0.5)
    dW = torch.zeros(cout, K, dtype=F32, device="cuda")
    ops.wgrad(x, cin, dy, cout, M, cout, K, conv_geom(B, H, W, cin, oh, ow, k, k, stride, pad), dW, K, rowscale=scale)
    ref_w = gw.permute(0, 2, 3, 1).reshape(cout, K) * scale[:, None]
    close("conv wgrad", dW, ref_w, 1e-2, 1e-2 * float(ref_w.abs().max()))


def test_stem_conv_and_wgrad(ops):
    B, H, W = 2, 37, 45
    oh, ow = _out(H, 7, 2, 3), _out(W, 7, 2, 3)
    img = dev(rnd(B, H, W, 3, seed=1))
    x4 = torch.zeros(B, H, W, 4, dtype=BF, device="cuda")
    ops.image_to_nhwc4(img, x4, B * H * W)
    assert torch.equal(x4[..., :3], img.to(BF)) and float(x4[..., 3].abs().max()) == 0
    w = dev(rnd(64, 7, 7, 3, scale=147 ** -0.5, seed=2).to(BF))
    wp = torch.zeros(64, 7, 8, 4, dtype=BF, device="cuda")
    wp[:, :, :7, :3] = w
    shift = dev(rnd(64, seed=3))
    y = torch.zeros(B, oh, ow, 64, dtype=BF, device="cuda")
    M = B * oh * ow
    ops.igemm(x4, wp, M, 64, 224, 4, 224, conv_geom(B, H, W, 4, oh, ow, 7, 8, 2, 3), bias=shift, relu=True, C=y, ldc=64)
    xt = img.to(BF).float().permute(0, 3, 1, 2)
    wt = w.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.conv2d(xt, wt, bias=shift, stride=2, padding=3)
    close("stem fwd", y, F.relu(ref).permute(0, 2, 3, 1), 1e-2, 2e-2)
    dy = dev(rnd(B, oh, ow, 64, seed=5).to(BF))
    gw, = torch.autograd.grad(ref, [wt], dy.float().permute(0, 3, 1, 2))
    dW = torch.zeros(64, 224, dtype=F32, device="cuda")
    ops.wgrad(x4, 4, dy, 64, M, 64, 224, conv_geom(B, H, W, 4, oh, ow, 7, 8, 2, 3), dW, 224)
    dWv = dW.view(64, 7, 8, 4)
    assert float(dWv[:, :, 7, :].abs().max()) == 0 and float(dWv[..., 3].abs().max()) == 0   # padding taps get no gradient
    ref_w = gw.permute(0, 2, 3, 1)
    close("stem wgrad", dWv[:, :, :7, :3], ref_w, 1e-2, 1e-2 * float(ref_w.abs().max()))


def test_linear_wgrad_with_padding_and_bias(ops):
    M, N, K = 1200, 92, 256
    x = dev(rnd(M, K, seed=1).to(BF))
    dy = torch.zeros(M, 96, dtype=BF, device="cuda")
    dy[:, :N] = dev(rnd(M, N, seed=2).to(BF))
    dW = torch.zeros(N, K, dtype=F32, device="cuda")
    db = torch.zeros(N, dtype=F32, device="cuda")
    ops.wgrad(x, K, dy, 96, M, N, K, ops.plain_geom(M, K), dW, K, dbias=db)
    close("linear wgrad", dW, dy[:, :N].float().t() @ x.float(), 1e-2, 0.05)
    close("bias grad", db, dy[:, :N].float().sum(0), 1e-3, 1e-2)
    # sub-block of a packed in_proj gradient: rows [256, 512) of a [768, 256] tensor
    dWp = torch.zeros(768, K, dtype=F32, device="cuda")
    dbp = torch.zeros(768, dtype=F32, device="cuda")
    dy2 = dev(rnd(M, 256, seed=3).to(BF))
    ops.wgrad(x, K, dy2, 256, M, 256, K, ops.plain_geom(M, K), dWp[256:], K, dbias=dbp[256:])
    close("packed wgrad", dWp[256:512], dy2.float().t() @ x.float(), 1e-2, 0.05)
    assert float(dWp[:256].abs().max()) == 0 and float(dWp[512:].abs().max()) == 0
    close("packed bias", dbp[256:512], dy2.float().sum(0), 1e-3, 1e-2)


# ------------------------------------------------------------------------------------------------ attention
def attn_ref(q, k, v, scale, keep=None, p=0.0):
    B, Lq, d = q.shape
    Lk, H = k.shape[1], d // 32
    qh = q.view(B, Lq, H, 32).transpose(1, 2)
    kh = k.view(B, Lk, H, 32).transpose(1, 2)
    vh = v.view(B, Lk, H, 32).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    w = torch.softmax(s, -1)
    if keep is not None:
        w = w * keep.view(B, H, Lq, Lk).float() / (1 - p)
    return (w @ vh).transpose(1, 2).reshape(B, Lq, d), torch.logsumexp(s, -1)


@pytest.mark.parametrize("B,Lq,Lk,drop", [(2, 100, 100, 0.0), (2, 100, 300, 0.0), (1, 300, 300, 0.0), (2, 70, 130, 0.1),
                                           (1, 1050, 1050, 0.0)])
def test_attention_fwd_bwd(ops, B, Lq, Lk, drop):
    H, d = 8, 256
    scale = 32 ** -0.5
    qk = dev(rnd(B * Lq, 2 * d, seed=1).to(BF))          # q lives in a [.,512] buffer like the engine's
    q = qk[:, :d]
    k = dev(rnd(B * Lk, d, seed=2).to(BF))
    v = dev(rnd(B * Lk, d, seed=3).to(BF))
    o = torch.zeros(B * Lq, d, dtype=BF, device="cuda")
    lse = torch.zeros(B * H * Lq, dtype=F32, device="cuda")
    seed_dev = torch.tensor([99], dtype=torch.int64, device="cuda")
    kw = dict(drop_p=drop, seed=3, site=11, seed_ptr=seed_dev) if drop > 0 else {}
    ops.attn_fwd(q, k, v, 2 * d, d, d, o, d, lse, B, H, Lq, Lk, scale, **kw)
    keep = None
    if drop > 0:
        keep = torch.zeros(B * H * Lq, Lk, dtype=torch.uint8, device="cuda")
        ops.attn_dropout_mask(keep, B * H * Lq, Lk, drop, 3, 11, seed_dev)
    qf = q.float().reshape(B, Lq, d).clone().requires_grad_(True)
    kf = k.float().view(B, Lk, d).clone().requires_grad_(True)
    vf = v.float().view(B, Lk, d).clone().requires_grad_(True)
    ref, ref_lse = attn_ref(qf, kf, vf, scale, keep, drop)
    close("attn out", o.view(B, Lq, d), ref, 2e-2, 2e-2)
    close("attn lse", lse.view(B, H, Lq), ref_lse, 1e-3, 2e-3)
    do = dev(rnd(B * Lq, d, seed=5).to(BF))
    gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], do.float().view(B, Lq, d))
    dqk = torch.zeros(B * Lq, 2 * d, dtype=BF, device="cuda")
    dk = torch.zeros(B * Lk, d, dtype=BF, device="cuda")
    dv = torch.zeros(B * Lk, d, dtype=BF, device="cuda")
    delta = torch.zeros(B * H * max(Lq, Lk), dtype=F32, device="cuda")
    ops.attn_bwd(q, k, v, o, do, 2 * d, d, d, d, d, lse, delta, dqk, dk, dv, 2 * d, d, d, B, H, Lq, Lk, scale, **kw)
    for name, got, ref_g in (("dq", dqk[:, :d], gq), ("dk", dk, gk), ("dv", dv, gv)):
        close("attn " + name, got.reshape(ref_g.shape), ref_g, 3e-2, 3e-2 * float(ref_g.abs().max()))
    assert float(dqk[:, d:].abs().max()) == 0


def test_attention_bwd_very_negative_scores(ops):
    """rows whose scores are all << 0 (lse < -88): exp(-lse) overflows on the zero-filled padding keys of the last tile;
    the gradients must stay finite (regression: inf * 0 in dQ)"""
    B, H, d, Lq, Lk, scale = 1, 8, 256, 70, 100, 32 ** -0.5
    q = dev((rnd(B * Lq, d, seed=1) * 0.1 - 10.0).to(BF))
    k = dev((rnd(B * Lk, d, seed=2) * 0.1 + 10.0).to(BF))
    v = dev(rnd(B * Lk, d, seed=3).to(BF))
    o = torch.zeros(B * Lq, d, dtype=BF, device="cuda")
    lse = torch.zeros(B * H * Lq, dtype=F32, device="cuda")
    ops.attn_fwd(q, k, v, d, d, d, o, d, lse, B, H, Lq, Lk, scale)
    assert float(lse.max()) < -300
    qf, kf, vf = (x.float().view(B, -1, d).clone().requires_grad_(True) for x in (q, k, v))
    ref, ref_lse = attn_ref(qf, kf, vf, scale, None, 0.0)
    close("attn out", o.view(B, Lq, d), ref, 2e-2, 2e-2)
    do = dev(rnd(B * Lq, d, seed=5).to(BF))
    gq, gk, gv = torch.autograd.grad(ref, [qf, kf, vf], do.float().view(B, Lq, d))
    dq, dk, dv = (torch.zeros(B * n, d, dtype=BF, device="cuda") for n in (Lq, Lk, Lk))
    delta = torch.zeros(B * H * max(Lq, Lk), dtype=F32, device="cuda")
    ops.attn_bwd(q, k, v, o, do, d, d, d, d, d, lse, delta, dq, dk, dv, d, d, d, B, H, Lq, Lk, scale)
    for name, got in (("dq", dq), ("dk", dk), ("dv", dv)):
        assert bool(torch.isfinite(got.float()).all()), name
    # (dQ / dK cancel catastrophically against the +-10 common mode of this contrived input; dV does not)
    close("attn dv", dv.reshape(gv.shape), gv, 5e-2, 5e-2 * float(gv.abs().max()))


# ------------------------------------------------------------------------------------------------ layer norm / elementwise
def test_layernorm_fwd_bwd(ops):
    M, S, d = 1000, 250, 256
    x = dev((rnd(M, d, seed=1) * 2 + 0.5).to(BF))
    g, b = dev(1 + 0.1 * rnd(d, seed=2)), dev(0.1 * rnd(d, seed=3))
    pos = dev(rnd(S, d, seed=4).to(BF))
    y, y2 = torch.zeros(M, d, dtype=BF, device="cuda"), torch.zeros(M, d, dtype=BF, device="cuda")
    mean, rstd = torch.zeros(M, device="cuda"), torch.zeros(M, device="cuda")
    ops.layernorm_fwd(x, g, b, y, y2, pos, S, mean, rstd, M)
    xf = x.float().clone().requires_grad_(True)
    gf, bf_ = g.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = F.layer_norm(xf, (d,), gf, bf_, 1e-5)
    close("ln y", y, ref, 1e-2, 1e-2)
    close("ln y2", y2, y.float() + pos.float().repeat(M // S, 1), 1e-2, 1e-2)
    close("ln mean", mean, x.float().mean(-1), 1e-4, 1e-4)
    dy, dy2 = dev(rnd(M, d, seed=5).to(BF)), dev(rnd(M, d, seed=6).to(BF))
    gx, gg, gb = torch.autograd.grad(ref, [xf, gf, bf_], dy.float() + dy2.float())
    dx, dxd = torch.zeros(M, d, dtype=BF, device="cuda"), torch.zeros(M, d, dtype=BF, device="cuda")
    dg, db = torch.zeros(d, device="cuda"), torch.zeros(d, device="cuda")
    seed_dev = torch.tensor([5], dtype=torch.int64, device="cuda")
    ops.layernorm_bwd(dy, dy2, x, g, mean, rstd, dx, dxd, 0.1, 9, 4, seed_dev, dg, db, M)
    close("ln dx", dx, gx, 2e-2, 2e-2)
    close("ln dgamma", dg, gg, 1e-2, 0.05)
    close("ln dbeta", db, gb, 1e-2, 0.05)
    keep = torch.zeros(M, d, dtype=torch.uint8, device="cuda")
    ops.dropout_mask(keep, M, d, 0.1, 9, 4, seed_dev)
    close("ln dx_drop", dxd, dx.float() * keep.float() / 0.9, 1e-2, 1e-2)


def test_elementwise_and_colsum(ops):
    M, S, d = 600, 150, 256
    x, pos = dev(rnd(M, d, seed=1).to(BF)), dev(rnd(S, d, seed=2).to(BF))
    out = torch.zeros(M, d, dtype=BF, device="cuda")
    ops.add_rowbcast(x, pos, out, M, S, d)
    close("add_rowbcast", out, x.float() + pos.float().repeat(M // S, 1), 1e-2, 1e-2)
    y = dev(rnd(M, d, seed=3).to(BF))
    ops.add(x, y, out, M * d)
    close("add", out, x.float() + y.float(), 1e-2, 1e-2)
    cs = torch.zeros(92, device="cuda")
    xx = dev(rnd(5000, 96, seed=4).to(BF))
    sc = dev(rnd(92, seed=5))
    ops.colsum(xx, 96, 5000, 92, sc, cs)
    close("colsum", cs, xx[:, :92].float().sum(0) * sc, 1e-3, 2e-2)


def test_maxpool_fwd_bwd(ops):
    B, H, W, C = 2, 21, 27, 64
    oh, ow = _out(H, 3, 2, 1), _out(W, 3, 2, 1)
    x = dev(F.relu(rnd(B, H, W, C, seed=1)).to(BF))
    y = torch.zeros(B, oh, ow, C, dtype=BF, device="cuda")
    arg = torch.zeros(B, oh, ow, C, dtype=torch.uint8, device="cuda")
    ops.maxpool_fwd(x, y, arg, B, H, W, C, oh, ow)
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(F.pad(xt, (1, 1, 1, 1)), 3, 2)
    assert torch.equal(y.float(), ref.permute(0, 2, 3, 1))
    dy = dev(rnd(B, oh, ow, C, seed=2).to(BF))
    # use strictly positive, tie-free inputs for the gradient comparison
    x2 = dev((torch.rand(B, H, W, C, generator=torch.Generator().manual_seed(3)) + 0.01).to(BF))
    ops.maxpool_fwd(x2, y, arg, B, H, W, C, oh, ow)
    xt = x2.float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(F.pad(xt, (1, 1, 1, 1)), 3, 2)
    gx, = torch.autograd.grad(ref, [xt], dy.float().permute(0, 3, 1, 2))
    dx = torch.zeros(B, H, W, C, dtype=BF, device="cuda")
    ops.maxpool_bwd(dy, arg, dx, B, H, W, C, oh, ow)
    # bf16 ties inside a window can route the gradient to another equal element: compare sums per window-safe tolerance
    rel = float((dx.float() - gx.permute(0, 2, 3, 1)).norm() / gx.norm())
    assert rel < 0.05, rel
    # the stem's ReLU mask travels in the argmax (tap 15 = non-positive maximum): gradient w.r.t. the PRE-activation
    x0 = rnd(B, H, W, C, seed=4) - 0.8                               # mostly negative: many all-zero windows after the ReLU
    x3 = dev(F.relu(x0).to(BF))
    ops.maxpool_fwd(x3, y, arg, B, H, W, C, oh, ow)
    x0t = x0.to(BF).float().permute(0, 3, 1, 2).requires_grad_(True)
    ref = F.max_pool2d(F.pad(F.relu(x0t), (1, 1, 1, 1)), 3, 2)
    assert torch.equal(y.float().cpu(), ref.permute(0, 2, 3, 1).detach().cpu())
    gx, = torch.autograd.grad(ref, [x0t], dy.float().cpu().permute(0, 3, 1, 2))
    ops.maxpool_bwd(dy, arg, dx, B, H, W, C, oh, ow)
    got = dx.float().cpu()
    assert float(got[(x3 <= 0).cpu()].abs().max()) == 0.0             # nothing flows where the stem output is zero
    rel = float((got - gx.permute(0, 2, 3, 1)).norm() / gx.norm())
    assert rel < 0.05, rel


@pytest.mark.parametrize("H,W", [(21, 27), (64, 90), (7, 8)])
def test_maxpool_fwd_argmax_tie_rule(ops, H, W):
    """The packed bf16x2 forward (maxpool_fwd_bf16_kernel) against a torch restatement of the rule the backward pass relies on: taps
    scanned in row-major order, padding absent, the FIRST maximum wins (strict >), a maximum that is not positive is stored as tap
    15 -- on inputs full of ties (values quantised to quarters, exact zeros after the ReLU)."""
    B, C = 2, 64
    oh, ow = _out(H, 3, 2, 1), _out(W, 3, 2, 1)
    g = torch.Generator().manual_seed(9)
    for x in (torch.round(F.relu(torch.randn(B, H, W, C, generator=g)) * 4) / 4, F.relu(torch.randn(B, H, W, C, generator=g) - 0.8),
              torch.randn(B, H, W, C, generator=g)):
        xb = dev(x.to(BF))
        y = torch.zeros(B, oh, ow, C, dtype=BF, device="cuda")
        arg = torch.full((B, oh, ow, C), 99, dtype=torch.uint8, device="cuda")
        ops.maxpool_fwd(xb, y, arg, B, H, W, C, oh, ow)
        torch.cuda.synchronize()
        xp = F.pad(xb.float().cpu().permute(0, 3, 1, 2), (1, 1, 1, 1), value=float("-inf"))
        patches = F.unfold(xp, 3, stride=2).view(B, C, 9, oh, ow)
        mx, ref_arg = patches.max(dim=2).values, patches.argmax(dim=2)         # argmax: index of the first maximal value
        ref_arg[~(mx > 0)] = 15
        assert torch.equal(y.float().cpu(), mx.permute(0, 2, 3, 1))
        assert torch.equal(arg.cpu().long(), ref_arg.permute(0, 2, 3, 1))


@pytest.mark.parametrize("H,W", [(21, 27), (64, 90), (7, 8)])
def test_maxpool_bwd_equals_sequential_fp32_scatter(ops, H, W):
    """The packed backward (maxpool_bwd_bf16_kernel): every window's gradient goes to the position its stored tap names (tap 15:
    nowhere), contributions to one input pixel are added in fp32 in window order (oy, then ox) and rounded to bf16 once -- checked
    bit for bit against numpy's sequential scatter-add, on tie-rich inputs (several windows often name the same pixel)."""
    B, C = 2, 64
    oh, ow = _out(H, 3, 2, 1), _out(W, 3, 2, 1)
    g = torch.Generator().manual_seed(11)
    x = dev((torch.round(F.relu(torch.randn(B, H, W, C, generator=g)) * 2) / 2).to(BF))
    y = torch.zeros(B, oh, ow, C, dtype=BF, device="cuda")
    arg = torch.zeros(B, oh, ow, C, dtype=torch.uint8, device="cuda")
    ops.maxpool_fwd(x, y, arg, B, H, W, C, oh, ow)
    dy = dev(torch.randn(B, oh, ow, C, generator=g).to(BF))
    dx = torch.full((B, H, W, C), 7.0, dtype=BF, device="cuda")
    ops.maxpool_bwd(dy, arg, dx, B, H, W, C, oh, ow)
    torch.cuda.synchronize()
    a = arg.cpu().numpy().astype(np.int64)
    bb, oy, ox, cc = np.meshgrid(np.arange(B), np.arange(oh), np.arange(ow), np.arange(C), indexing="ij")
    keep = a != 15
    iy, ix = 2 * oy - 1 + a // 3, 2 * ox - 1 + a % 3
    assert keep.any() and (~keep).any()
    assert ((iy[keep] >= 0) & (iy[keep] < H) & (ix[keep] >= 0) & (ix[keep] < W)).all()
    ref = np.zeros((B, H, W, C), dtype=np.float32)
    np.add.at(ref, (bb[keep], iy[keep], ix[keep], cc[keep]), dy.float().cpu().numpy()[keep])      # unbuffered: sequential, C order
    assert torch.equal(dx.cpu(), torch.from_numpy(ref).to(BF))


# ------------------------------------------------------------------------------------------------ matcher / loss
def _run_matcher(ops, logits, boxes, t_bbox, t_class, want_cost=True):
    P, Q, C = logits.shape
    B = t_bbox.shape[0]
    out = dict(p_indices=torch.zeros(P, Q, dtype=torch.int64, device="cuda"), t_indices=torch.zeros(P, Q, dtype=torch.int64, device="cuda"),
               p_selector=torch.zeros(P, Q, dtype=torch.uint8, device="cuda"), match=torch.zeros(P, Q, dtype=torch.int32, device="cuda"),
               status=torch.zeros(P, dtype=torch.int32, device="cuda"),
               cost=torch.zeros(P, Q, 100, dtype=F32, device="cuda") if want_cost else None)
    ops.matcher(dev(logits), C, dev(boxes), dev(t_bbox), dev(t_class), P, B, Q, C, out["p_indices"], out["t_indices"],
                out["p_selector"], out["match"], out["cost"], out["status"])
    torch.cuda.synchronize()
    return {k: (v.cpu() if v is not None else None) for k, v in out.items()}


def test_matcher_vs_reference_golden(ops, golden):
    g = golden
    logits, boxes = torch.from_numpy(g["m_logits"]), torch.from_numpy(g["m_boxes"])
    tb, tc = torch.from_numpy(g["m_t_bbox"]), torch.from_numpy(g["m_t_class"])
    r = _run_matcher(ops, logits, boxes, tb, tc)
    assert int(r["status"].abs().sum()) == 0
    for b in range(logits.shape[0]):
        n = int(tb[b, 0, 0])
        np.testing.assert_allclose(r["cost"][b, :, :n].numpy(), g[f"m_cost_{b}"], rtol=1e-5, atol=3e-6)
        assert np.array_equal(r["p_indices"][b, :n].numpy(), g[f"m_p_indices_{b}"])          # bit-exact vs the reference
        assert np.array_equal(r["t_indices"][b, :n].numpy(), g[f"m_t_indices_{b}"])
        assert np.array_equal(r["p_selector"][b].numpy().astype(bool), g[f"m_p_selector_{b}"])
        assert int((r["p_indices"][b, n:] != -1).sum()) == 0


def test_matcher_c5_vs_scipy_on_own_cost(ops):
    """BASELINE config 5: 100 queries x 20 targets x batch 256; indices must equal scipy's on the kernel's own cost."""
    from scipy.optimize import linear_sum_assignment
    from oracle import detr_oracle as O
    B, Q, C = 256, 100, 92
    g = torch.Generator().manual_seed(1234)
    logits = torch.randn(B, Q, C, generator=g)
    boxes = torch.cat([torch.rand(B, Q, 2, generator=g) * 0.9 + 0.05, torch.rand(B, Q, 2, generator=g) * 0.48 + 0.02], -1)
    tb, tc = O.synthetic_targets(B, n=20, seed=1234)
    r = _run_matcher(ops, logits, boxes, tb, tc)
    assert int(r["status"].abs().sum()) == 0
    n_oracle_equal = 0
    for b in range(B):
        rows, cols = linear_sum_assignment(r["cost"][b, :, :20].numpy())
        assert np.array_equal(r["p_indices"][b, :20].numpy(), rows) and np.array_equal(r["t_indices"][b, :20].numpy(), cols)
        ti, pi, _, _, _, _ = O.hungarian_matching(tb[b], tc[b], boxes[b], logits[b])
        n_oracle_equal += int(np.array_equal(pi.numpy(), rows) and np.array_equal(ti.numpy(), cols))
    assert n_oracle_equal == B         # and equal to the oracle's assignment from the same inputs


def test_matcher_ties_and_edge_cases(ops):
    """tie-heavy problems (identical queries / duplicated targets), n = 0, n = 99, NaN detection"""
    from scipy.optimize import linear_sum_assignment
    from oracle import detr_oracle as O
    B, Q, C = 6, 100, 92
    g = torch.Generator().manual_seed(7)
    logits = torch.zeros(B, Q, C)                         # all queries identical class scores
    boxes = torch.cat([torch.rand(B, 1, 2, generator=g).expand(B, Q, 2) * 0.5 + 0.25, torch.full((B, Q, 2), 0.2)], -1).contiguous()
    boxes[:, ::3, 0] += 0.125                             # three groups of exactly tied queries
    tb, tc = O.synthetic_targets(B, n=7, seed=3)
    tb[1, 2] = tb[1, 1]                                   # duplicated target
    tb[2, 0, 0] = 0                                       # image without targets
    tb3, tc3 = O.synthetic_targets(1, n=99, seed=5)
    tb[3], tc[3] = tb3[0], tc3[0]
    r = _run_matcher(ops, logits, boxes, tb, tc)
    assert int(r["status"].abs().sum()) == 0
    for b in range(B):
        n = int(tb[b, 0, 0])
        rows, cols = linear_sum_assignment(r["cost"][b, :, :n].numpy())
        assert np.array_equal(r["p_indices"][b, :n].numpy(), rows), b
        assert np.array_equal(r["t_indices"][b, :n].numpy(), cols), b
        assert int((r["match"][b] >= 0).sum()) == n
    logits[0, 5, 3] = float("nan")
    r = _run_matcher(ops, logits, boxes, tb, tc)
    assert int(r["status"][0]) == 1 and int((r["match"][0] >= 0).sum()) == 0 and int(r["status"][1:].abs().sum()) == 0


def test_set_loss_vs_reference_golden_and_grads(ops, golden):
    from oracle import detr_oracle as O
    g = golden
    logits, boxes = torch.from_numpy(g["l_logits"]), torch.from_numpy(g["l_boxes"])       # [6,B,Q,C]
    tb, tc = torch.from_numpy(g["l_t_bbox"]), torch.from_numpy(g["l_t_class"])
    L, B, Q, C = logits.shape
    r = _run_matcher(ops, logits.reshape(L * B, Q, C), boxes.reshape(L * B, Q, 4), tb, tc, want_cost=False)
    sums = torch.zeros(L, 8, device="cuda")
    losses = torch.zeros(L, 6, device="cuda")
    total = torch.zeros(1, device="cuda")
    dl = torch.zeros(L * B * Q, 96, dtype=BF, device="cuda")
    db = torch.zeros(L * B * Q, 32, dtype=BF, device="cuda")
    lg_d, bx_d = dev(logits.reshape(-1, C)), dev(boxes.reshape(-1, 4))
    ops.set_loss(lg_d, C, bx_d, dev(tb), dev(tc), dev(r["match"]), L, B, Q, C, 91, None, 1.0, sums, losses, total, dl, 96, db, 32)
    torch.cuda.synchronize()
    names = ("label_cost", "true_neg", "true_pos", "pos_accuracy", "giou_loss", "l1_loss")
    ref = dict(zip([str(k) for k in g["l_keys"]], g["l_values"]))
    for l in range(L):
        suf = "" if l == L - 1 else f"_{l}"
        for k, nme in enumerate(names):
            assert abs(float(losses[l, k]) - float(ref[nme + suf])) < 1e-4 + 1e-4 * abs(float(ref[nme + suf])), (nme + suf)
    assert abs(float(total) - float(g["l_total"])) < 2e-4 * float(g["l_total"])
    # gradients vs autograd of the oracle under the same assignment
    lg = logits.clone().requires_grad_(True)
    bx = boxes.clone().requires_grad_(True)
    out = {"pred_logits": lg[L - 1], "pred_boxes": bx[L - 1], "aux": [{"pred_logits": lg[i], "pred_boxes": bx[i]} for i in range(L - 1)]}
    tot, _ = O.get_losses(out, tb, tc, 91, r["match"].view(L, B, Q))
    gl, gb = torch.autograd.grad(tot, [lg, bx])
    close("d_logits", dl[:, :C].view(L, B, Q, C), gl, 1e-2, 1e-2 * float(gl.abs().max()))
    assert float(dl[:, C:].abs().max()) == 0
    gpre = gb * boxes * (1 - boxes)
    close("d_boxpre", db[:, :4].view(L, B, Q, 4), gpre, 2e-2, 1e-2 * float(gpre.abs().max()))
    assert float(db[:, 4:].abs().max()) == 0


# ------------------------------------------------------------------------------------------------ optimizer
def test_adam_clipnorm_and_prep_weight(ops):
    from oracle import detr_oracle as O
    sizes = [5000, 64, 300000, 7, 1024]
    offs, off = [], 0
    for n in sizes:
        offs.append(off)
        off = (off + n + 63) // 64 * 64
    total = off
    g = torch.Generator().manual_seed(0)
    P = torch.randn(total, generator=g)
    G = torch.zeros(total)
    for (o, n), sc in zip(zip(offs, sizes), (1e-3, 10.0, 1e-2, 5.0, 1e-5)):      # norms below and above clipnorm
        G[o:o + n] = torch.randn(n, generator=g) * sc
    table = torch.tensor([[o, n] for o, n in zip(offs, sizes)], dtype=torch.int64)
    grp = torch.tensor([0, 0, 1, 1, 2], dtype=torch.int32)
    lrs = torch.zeros(8)
    lrs[:3] = torch.tensor([1e-2, 1e-3, 5e-2])
    en = torch.zeros(8, dtype=torch.uint8)
    en[:3] = torch.tensor([1, 1, 0], dtype=torch.uint8)
    p_d, g_d = dev(P.clone()), dev(G.clone())
    m_d, v_d = torch.zeros(total, device="cuda"), torch.zeros(total, device="cuda")
    steps, norms = torch.zeros(8, dtype=torch.int32, device="cuda"), torch.zeros(len(sizes), device="cuda")
    ref_p = [P[o:o + n].clone() for o, n in zip(offs, sizes)]
    ref_m = [torch.zeros(n) for n in sizes]
    ref_v = [torch.zeros(n) for n in sizes]
    for step in (1, 2, 3):
        ops.adam_clipnorm(p_d, g_d, m_d, v_d, dev(table), dev(grp), dev(lrs), dev(en), len(sizes), total, 0.1, steps, norms)
        for t, (o, n) in enumerate(zip(offs, sizes)):
            if en[int(grp[t])]:
                O.adam_clipnorm_step(ref_p[t], G[o:o + n].clone(), ref_m[t], ref_v[t], step, float(lrs[int(grp[t])]), 0.1)
    torch.cuda.synchronize()
    assert steps.cpu()[:3].tolist() == [3, 3, 0]
    for t, (o, n) in enumerate(zip(offs, sizes)):
        close(f"adam tensor {t}", p_d[o:o + n], ref_p[t], 1e-5, 1e-6)
    close("norms", norms.cpu().sqrt(), torch.stack([G[o:o + n].norm() for o, n in zip(offs, sizes)]), 1e-4, 1e-6)
    # chunked / vectorised variant must give the same result
    CH = 8192
    chunks = torch.tensor([[t, o + c, min(CH, n - c)] for t, (o, n) in enumerate(zip(offs, sizes)) for c in range(0, n, CH)], dtype=torch.int32)
    p2, m2, v2 = dev(P.clone()), torch.zeros(total, device="cuda"), torch.zeros(total, device="cuda")
    steps2, norms2 = torch.zeros(8, dtype=torch.int32, device="cuda"), torch.zeros(len(sizes), device="cuda")
    for step in (1, 2, 3):
        ops.adam_clipnorm_chunked(p2, g_d, m2, v2, dev(chunks), chunks.shape[0], dev(grp), dev(lrs), dev(en), len(sizes), 0.1, steps2, norms2)
    torch.cuda.synchronize()
    close("chunked adam params", p2, p_d, 1e-6, 1e-7)
    close("chunked adam v", v2, v_d, 1e-5, 1e-12)
    assert steps2.cpu()[:3].tolist() == [3, 3, 0]
    # one step issued as two calls over disjoint chunk ranges (whole variables each; the engine applies the stem kernel last)
    p3, m3, v3 = dev(P.clone()), torch.zeros(total, device="cuda"), torch.zeros(total, device="cuda")
    steps3, norms3 = torch.zeros(8, dtype=torch.int32, device="cuda"), torch.zeros(len(sizes), device="cuda")
    n0 = int((chunks[:, 0] == 0).sum())
    ch_d = dev(chunks)
    for step in (1, 2, 3):
        ops.adam_clipnorm_chunked(p3, g_d, m3, v3, ch_d, chunks.shape[0] - n0, dev(grp), dev(lrs), dev(en), len(sizes), 0.1, steps3, norms3,
                                  first_chunk=n0, prologue=True)
        ops.adam_clipnorm_chunked(p3, g_d, m3, v3, ch_d, n0, dev(grp), dev(lrs), dev(en), len(sizes), 0.1, steps3, norms3,
                                  first_chunk=0, prologue=False)
    torch.cuda.synchronize()
    # (the per-variable norms are sums of fp32 atomics: equal up to the summation order)
    close("split adam params", p3, p2, 1e-6, 1e-7)
    close("split adam v", v3, v2, 1e-5, 1e-12)
    assert steps3.cpu()[:3].tolist() == [3, 3, 0]
    # prep_weight
    N, taps, Cin = 96, 9, 64
    master = dev(rnd(N, taps, Cin, seed=4))
    fold = dev(rnd(N, seed=5))
    Wf = torch.zeros(N, taps * Cin, dtype=BF, device="cuda")
    Wd = torch.zeros(Cin, taps, 128, dtype=BF, device="cuda")
    ops.prep_weight(master, fold, N, taps, Cin, Wf, taps * Cin, Wd, 128)
    refw = (master * fold[:, None, None]).to(BF)
    assert torch.equal(Wf.view(N, taps, Cin), refw)
    assert torch.equal(Wd[:, :, :N], refw.permute(2, 1, 0)) and float(Wd[:, :, N:].abs().max()) == 0
