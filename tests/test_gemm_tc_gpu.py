"""tcgen05 / TMA / TMEM GEMM (csrc/gemm_tc.cu) against a PyTorch fp32 reference and against the mma.sync kernel."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
BF, F32 = torch.bfloat16, torch.float32


@pytest.fixture(scope="module", params=["one-tile-per-CTA", "persistent", "streaming"])
def ops(request):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import _lib, ops as o
    _lib.check(_lib.lib().detrb_check_device())
    old = o.set_tc_persistent(2 if request.param == "persistent" else 0)      # 2: persistent kernel wherever supported
    olds = o.set_tc_stream(2 if request.param == "streaming" else 0)          # 2: streaming kernel wherever supported (bn = 0 launches)
    yield o
    o.set_tc_persistent(old)
    o.set_tc_stream(olds)


def rnd(*shape, scale=1.0, seed=0):
    return (torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()


def check(name, got, ref, rtol, atol):
    got, ref = got.float(), ref.float()
    err = (got - ref).abs()
    bad = err > atol + rtol * ref.abs()
    assert int(bad.sum()) == 0, f"{name}: {int(bad.sum())}/{bad.numel()} bad, max err {float(err.max()):.4g}, " \
                                f"rel {float((got - ref).norm() / ref.norm()):.4g}, first bad {torch.nonzero(bad)[0].tolist()}"


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (128, 128, 64, 128), (256, 128, 128, 128), (1050, 256, 256, 128),
                                      (1050, 256, 256, 64), (4175, 64, 256, 64), (300, 2048, 256, 128), (777, 256, 2048, 128),
                                      (100, 256, 256, 0), (8400, 512, 1024, 0), (130, 72, 64, 64)])
def test_gemm_tc_plain(ops, M, N, K, bn):
    A = rnd(M, K, seed=1).to(BF)
    W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
    bias = rnd(N, seed=3)
    C = torch.full((M, N), 7.0, dtype=BF, device="cuda")
    Cf = torch.full((M, N), 7.0, dtype=F32, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, C=C, ldc=N, Cf=Cf, ldcf=N, force_tc=bn)
    torch.cuda.synchronize()
    ref = A.float() @ W.float().t() + bias
    check("tc f32", Cf, ref, 1e-3, 1e-3)
    check("tc bf16", C, ref, 1e-2, 1e-2)


@pytest.mark.parametrize("M,N,K,bn", [(128, 64, 64, 64), (1050, 256, 256, 128), (4175, 64, 256, 64), (130, 72, 64, 64), (8400, 2048, 256, 0),
                                      (777, 256, 2048, 0)])
def test_gemm_tc_tma_epilogue(ops, M, N, K, bn):
    """bf16-only output -> TMA-store epilogue with TMA-loaded residual and mask tiles"""
    A = rnd(M, K, seed=1).to(BF)
    W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
    bias, res, mask = rnd(N, seed=3), rnd(M, N, seed=4).to(BF), rnd(M, N, seed=5).to(BF)
    base = A.float() @ W.float().t() + bias
    C = torch.full((M + 3, N), 7.0, dtype=BF, device="cuda")          # 3 guard rows: the store must clip at M
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, relu=True, C=C, ldc=N, force_tc=bn)
    torch.cuda.synchronize()
    check("tma epi res+relu", C[:M], F.relu(base + res.float()), 1e-2, 1e-2)
    assert float((C[M:].float() - 7.0).abs().max()) == 0
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, mask=mask, ldm=N, mask_scale=1.5, C=C, ldc=N,
              force_tc=bn)
    torch.cuda.synchronize()
    check("tma epi res+mask", C[:M], torch.where(mask.float() > 0, (base + res.float()) * 1.5, torch.zeros_like(base)), 1e-2, 1e-2)
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), C=C, ldc=N, force_tc=bn)
    torch.cuda.synchronize()
    check("tma epi plain", C[:M], A.float() @ W.float().t(), 1e-2, 1e-2)


@pytest.mark.parametrize("M,N,K,bn", [(1050, 256, 512, 256), (8400, 512, 320, 256), (33600, 256, 2304, 256), (4175, 192, 64, 64), (20000, 320, 128, 128),
                                      (20000, 320, 128, 256), (128 * 148 * 3 + 5, 64, 64, 64)])
def test_gemm_tc_persistent_tiles(ops, M, N, K, bn):
    """persistent kernel: 256-wide tiles, ragged N (partial last tile / chunk), residual+mask ring of depth 1..4, many tiles per CTA"""
    if ops.set_tc_persistent(2) != 2:
        ops.set_tc_persistent(0)
        pytest.skip("persistent mode only")
    A = rnd(M, K, seed=1).to(BF)
    W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
    bias, res, mask = rnd(N, seed=3), rnd(M, N, seed=4).to(BF), rnd(M, N, seed=5).to(BF)
    base = A.float() @ W.float().t() + bias
    C = torch.full((M + 3, N), 7.0, dtype=BF, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, C=C, ldc=N, force_tc=bn)
    torch.cuda.synchronize()
    check("plain", C[:M], F.relu(base), 1e-2, 1e-2)
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, relu=True, C=C, ldc=N, force_tc=bn)
    torch.cuda.synchronize()
    check("res", C[:M], F.relu(base + res.float()), 1e-2, 1e-2)
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, mask=mask, ldm=N, mask_scale=1.5, C=C, ldc=N,
              force_tc=bn)
    torch.cuda.synchronize()
    check("res+mask", C[:M], torch.where(mask.float() > 0, (base + res.float()) * 1.5, torch.zeros_like(base)), 1e-2, 1e-2)
    assert float((C[M:].float() - 7.0).abs().max()) == 0


def test_gemm_tc_epilogues_match_mma_sync_kernel(ops):
    M, N, K = 1000, 256, 192
    A = rnd(M, 2 * K, seed=1).to(BF)[:, K:]            # strided A (lda = 2K)
    W = rnd(3 * N, K, scale=K ** -0.5, seed=2).to(BF)[N:2 * N]
    bias, res, mask = rnd(N, seed=3), rnd(M, N, seed=4).to(BF), rnd(M, N, seed=5).to(BF)
    seed_dev = torch.tensor([4242], dtype=torch.int64, device="cuda")
    cases = [dict(bias=bias, residual=res, ldr=N, relu=True), dict(bias=bias, mask=mask, ldm=N, mask_scale=1.0 / 0.9),
             dict(bias=bias, sigmoid=True), dict(bias=bias, residual=res, ldr=N, drop_p=0.1, seed=5, site=3, seed_ptr=seed_dev),
             dict(bias=bias, relu=True, drop_p=0.1, seed=5, site=4, seed_ptr=seed_dev)]
    for i, kw in enumerate(cases):
        C1, C2 = torch.zeros(M, N, dtype=BF, device="cuda"), torch.zeros(M, N, dtype=BF, device="cuda")
        ops.igemm(A, W, M, N, K, 2 * K, K, ops.plain_geom(M, K), C=C1, ldc=N, **kw)
        ops.igemm(A, W, M, N, K, 2 * K, K, ops.plain_geom(M, K), C=C2, ldc=N, force_tc=0, **kw)
        torch.cuda.synchronize()
        check(f"case {i}", C2, C1, 1e-2, 1e-2)
    # accumulate + strided scatter (1x1 stride-2 data gradient)
    B, oh, ow, H, Wd, cin, cout = 2, 7, 10, 13, 19, 256, 512
    dy = rnd(B * oh * ow, cout, seed=6).to(BF)
    wd = rnd(cin, cout, scale=cout ** -0.5, seed=7).to(BF)
    base = rnd(B * H * Wd, cin, seed=8).to(BF)
    xm = rnd(B * H * Wd, cin, seed=9).to(BF)
    g = dict(batch=B, IH=oh, IW=ow, Cin=cout, OH=oh, OW=ow, KH=1, KW=1, stride=1, pad=0, mode=0)
    outs = []
    for tc in (None, 0):
        dx = base.clone()
        ops.igemm(dy, wd, B * oh * ow, cin, cout, cout, cout, g, mask=xm, ldm=cin, C=dx, ldc=cin, out_stride=2, SH=H, SW=Wd,
                  accumulate=True, force_tc=tc)
        outs.append(dx)
    torch.cuda.synchronize()
    check("scatter-accumulate", outs[1], outs[0], 1e-2, 2e-2)


def test_engine_forward_with_tc_enabled(ops):
    import detr_tensorflow_b200 as D
    from oracle import detr_oracle as O
    P = O.init_params(seed=1)
    img = torch.randn(2, 160, 224, 3, generator=torch.Generator().manual_seed(1))
    cfg = D.TrainingConfig()
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0)
    old = ops.set_tc(1)
    try:
        out = model(img, training=False)
        torch.cuda.synchronize()
    finally:
        ops.set_tc(old)
    with torch.no_grad():
        ref = O.detr_forward(P, img)
    rel = float((out["pred_logits"].cpu() - ref["pred_logits"]).norm() / ref["pred_logits"].norm())
    assert rel < 5e-2, rel


# ------------------------------------------------------------------------------------------------ TMA-im2col convolutions
def conv_geom(B, ih, iw, cin, oh, ow, kh, kw, stride, pad, mode=0):
    return dict(batch=B, IH=ih, IW=iw, Cin=cin, OH=oh, OW=ow, KH=kh, KW=kw, stride=stride, pad=pad, mode=mode)


def _out(n, k, s, p):
    return (n + 2 * p - k) // s + 1


@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W,bn", [(64, 64, 3, 1, 1, 13, 19, 64), (128, 128, 3, 1, 1, 25, 42, 128),
                                                           (128, 128, 3, 2, 1, 13, 19, 0), (128, 128, 3, 2, 1, 12, 18, 0), (256, 256, 3, 2, 1, 100, 167, 0), (256, 512, 1, 2, 0, 13, 19, 0),
                                                           (256, 256, 3, 1, 1, 50, 84, 0), (64, 64, 3, 1, 1, 200, 334, 0)])
def test_conv_tc_forward_and_dgrad(ops, cin, cout, k, stride, pad, H, W, bn):
    B = 2
    oh, ow = _out(H, k, stride, pad), _out(W, k, stride, pad)
    x = rnd(B, H, W, cin, seed=1).to(BF)
    w = rnd(cout, k, k, cin, scale=(k * k * cin) ** -0.5, seed=2).to(BF)
    shift = rnd(cout, seed=3)
    res = rnd(B, oh, ow, cout, seed=4).to(BF)
    M, K = B * oh * ow, k * k * cin
    y_tc = torch.zeros(B, oh, ow, cout, dtype=BF, device="cuda")
    g = conv_geom(B, H, W, cin, oh, ow, k, k, stride, pad)
    ops.igemm(x, w, M, cout, K, cin, K, g, bias=shift, residual=res, ldr=cout, relu=True, C=y_tc, ldc=cout, force_tc=bn)
    torch.cuda.synchronize()
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias=shift, stride=stride, padding=pad)
    ref = F.relu(ref.permute(0, 2, 3, 1) + res.float())
    check("conv tc fwd", y_tc, ref, 1e-2, 2e-2)
    if stride == 1 or (stride == 2 and k == 3 and pad == 1):
        # data gradient: stride 1 -> flipped-tap convolution; 3x3 stride 2 -> four parity-class sub-convolutions
        dy = rnd(B, oh, ow, cout, seed=5).to(BF)
        wd = w.reshape(cout, k * k, cin).permute(2, 1, 0).contiguous()
        mask = rnd(B, H, W, cin, seed=6).to(BF)
        dx_a = torch.zeros(B, H, W, cin, dtype=BF, device="cuda")
        dx_b = torch.zeros_like(dx_a)
        gd = conv_geom(B, oh, ow, cout, H, W, k, k, stride, pad, mode=1)
        ops.igemm(dy, wd, B * H * W, cin, k * k * cout, cout, k * k * cout, gd, mask=mask, ldm=cin, C=dx_a, ldc=cin)
        ops.igemm(dy, wd, B * H * W, cin, k * k * cout, cout, k * k * cout, gd, mask=mask, ldm=cin, C=dx_b, ldc=cin, force_tc=bn)
        torch.cuda.synchronize()
        xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
        o = F.conv2d(xt, w.float().permute(0, 3, 1, 2), stride=stride, padding=pad)
        gx, = torch.autograd.grad(o, [xt], dy.float().permute(0, 3, 1, 2))
        refd = gx.permute(0, 2, 3, 1) * (mask.float() > 0)
        check("conv tc dgrad vs torch", dx_b, refd, 1e-2, 2e-2)
        check("conv tc dgrad vs mma.sync", dx_b, dx_a, 1e-2, 2e-2)


# ------------------------------------------------------------------------------------------------ tcgen05 weight gradient
@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W", [(64, 64, 1, 1, 0, 13, 19), (64, 64, 3, 1, 1, 13, 19), (128, 128, 3, 2, 1, 13, 19),
                                                       (256, 512, 1, 2, 0, 13, 19), (64, 256, 1, 1, 0, 9, 11), (256, 256, 3, 1, 1, 50, 84),
                                                       (256, 92, 1, 1, 0, 30, 40), (2048, 256, 1, 1, 0, 25, 42),
                                                       # >= 148 x 64 pixels and <= 64 output channels: wgrad_narrow_kernel (one CTA owns all of K)
                                                       (64, 64, 3, 1, 1, 100, 167), (256, 64, 1, 1, 0, 90, 131), (64, 48, 3, 1, 1, 77, 99),
                                                       (128, 64, 1, 1, 0, 80, 80), (64, 64, 1, 1, 0, 200, 334)])
def test_wgrad_tc(ops, cin, cout, k, stride, pad, H, W):
    B = 2
    oh, ow = _out(H, k, stride, pad), _out(W, k, stride, pad)
    M, K = B * oh * ow, k * k * cin
    x = rnd(B, H, W, cin, seed=1).to(BF)
    ldy = (cout + 7) // 8 * 8
    dy = torch.zeros(M, ldy, dtype=BF, device="cuda")
    dy[:, :cout] = rnd(M, cout, seed=2).to(BF)
    scale = rnd(cout, seed=3).abs() + 0.5
    g = conv_geom(B, H, W, cin, oh, ow, k, k, stride, pad)
    dW_a = torch.zeros(cout, K, dtype=F32, device="cuda")
    dW_b = torch.zeros(cout, K, dtype=F32, device="cuda")
    ops.wgrad(x, cin, dy, ldy, M, cout, K, g, dW_a, K, rowscale=scale)
    db = torch.zeros(cout, dtype=F32, device="cuda")
    ops.wgrad(x, cin, dy, ldy, M, cout, K, g, dW_b, K, rowscale=scale, dbias=db, force_tc=True)
    torch.cuda.synchronize()
    ref_b = dy[:, :cout].float().sum(0) * scale
    check("fused bias gradient", db, ref_b, 2e-3, 2e-3 * float(ref_b.abs().max()) + 1e-3)
    xt = x.float().permute(0, 3, 1, 2)
    wt = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    o = F.conv2d(xt, wt, stride=stride, padding=pad)
    gw, = torch.autograd.grad(o, [wt], dy[:, :cout].float().view(B, oh, ow, cout).permute(0, 3, 1, 2))
    ref = gw.permute(0, 2, 3, 1).reshape(cout, K) * scale[:, None]
    tol = 1e-2 * float(ref.abs().max())
    check("wgrad tc vs torch", dW_b, ref, 1e-2, tol)
    check("wgrad tc vs mma.sync", dW_b, dW_a, 1e-2, tol)


@pytest.mark.parametrize("tile", [2, 3, 4])
@pytest.mark.parametrize("cin,cout,k,stride,pad,H,W", [(256, 256, 3, 1, 1, 50, 84), (128, 128, 3, 1, 1, 31, 45), (256, 512, 1, 2, 0, 26, 38),
                                                       (1024, 256, 1, 1, 0, 25, 42), (256, 1024, 1, 1, 0, 25, 42), (512, 512, 3, 2, 1, 25, 42),
                                                       (192, 320, 3, 1, 1, 17, 23), (128, 512, 1, 1, 0, 50, 67)])
def test_wgrad_tc_wide_tiles(ops, tile, cin, cout, k, stride, pad, H, W):
    """The wider tiles of the tcgen05 weight-gradient kernel (wgrad_tc_kernel<NA, KT>: 128 x 256, 256 x 128, 256 x 256 out channels x
    k columns, one CTA per SM) against torch and against the base 128 x 128 tile; shapes with partial last tiles in both directions
    (K = 1152 / 1728 with KT = 256, N = 320 with NA = 2) and the fused bias gradient of the second accumulator."""
    B = 2
    oh, ow = _out(H, k, stride, pad), _out(W, k, stride, pad)
    M, K = B * oh * ow, k * k * cin
    x = rnd(B, H, W, cin, seed=1).to(BF)
    dy = rnd(M, cout, seed=2).to(BF)
    scale = rnd(cout, seed=3).abs() + 0.5
    g = conv_geom(B, H, W, cin, oh, ow, k, k, stride, pad)
    res = {}
    for mode in (1, tile):
        old = ops.set_wgrad_tile(mode)
        try:
            dW = torch.zeros(cout, K, dtype=F32, device="cuda")
            db = torch.zeros(cout, dtype=F32, device="cuda")
            ops.wgrad(x, cin, dy, cout, M, cout, K, g, dW, K, rowscale=scale, dbias=db, force_tc=True)
            torch.cuda.synchronize()
        finally:
            ops.set_wgrad_tile(old)
        res[mode] = (dW, db)
    xt = x.float().permute(0, 3, 1, 2)
    wt = torch.zeros(cout, cin, k, k, device="cuda", requires_grad=True)
    o = F.conv2d(xt, wt, stride=stride, padding=pad)
    gw, = torch.autograd.grad(o, [wt], dy.float().view(B, oh, ow, cout).permute(0, 3, 1, 2))
    ref = gw.permute(0, 2, 3, 1).reshape(cout, K) * scale[:, None]
    ref_b = dy.float().sum(0) * scale
    tol = 1e-2 * float(ref.abs().max())
    check("wide tile vs torch", res[tile][0], ref, 1e-2, tol)
    check("wide tile vs base tile", res[tile][0], res[1][0], 1e-3, 1e-3 * float(ref.abs().max()))
    check("wide tile bias gradient", res[tile][1], ref_b, 2e-3, 2e-3 * float(ref_b.abs().max()) + 1e-3)


# ------------------------------------------------------------------------------------------------ stem as a sliding-window GEMM
@pytest.mark.parametrize("H,W", [(37, 45), (64, 96), (160, 224), (400, 667)])
def test_stem_sliding_window_gemm_pool_and_wgrad(ops, H, W):
    """The production stem: zero-padded space-to-depth image [B, HP, WP, 16] read as a GEMM operand whose rows are overlapping
    128-byte windows (detrb_igemm_t.a_kb_rows = WP, lda = 16) == 7x7 / stride 2 / pad 3 convolution (resnet_backbone.py:11-12,
    20-24); max-pool forward / backward on the [HP, WP]-pitched output; weight gradient with the 7x7x3 column mask."""
    from detr_tensorflow_b200.engine import Engine
    B = 2
    H2, W2 = (H + 1) // 2, (W + 1) // 2
    HP, WP = H2 + 3, W2 + 3
    M = B * HP * WP
    img = rnd(B, H, W, 3, seed=1)
    s2d = torch.zeros(M + 3 * WP + 8, 16, dtype=BF, device="cuda")
    s2d[:M].fill_(7.0)                                     # the kernel must overwrite the padding with zeros
    ops.image_to_s2d16(img, s2d, B, H, W, 2, 2, HP, WP)
    dense = torch.zeros(B, H2, W2, 16, dtype=BF, device="cuda")
    ops.image_to_s2d16(img, dense, B, H, W)
    ref_pad = torch.zeros(B, HP, WP, 16, dtype=BF, device="cuda")
    ref_pad[:, 2:2 + H2, 2:2 + W2] = dense
    assert torch.equal(s2d[:M].view(B, HP, WP, 16), ref_pad)
    w7 = rnd(64, 7, 7, 3, scale=147 ** -0.5, seed=2).to(BF)
    w16 = Engine._stem_to_s2d(None, w7.float().cpu()).to(BF).cuda().reshape(64, 256).contiguous()
    shift = rnd(64, seed=3)
    y = torch.zeros(B, HP, WP, 64, dtype=BF, device="cuda")
    g = dict(batch=1, IH=1, IW=M, Cin=256, OH=1, OW=M, KH=1, KW=1, stride=1, pad=0, mode=0)
    ops.igemm(s2d, w16, M, 64, 256, 16, 256, g, bias=shift, relu=True, C=y, ldc=64, a_kb_rows=WP)
    torch.cuda.synchronize()
    xt = img.to(BF).float().permute(0, 3, 1, 2)
    wt = w7.float().permute(0, 3, 1, 2).requires_grad_(True)
    conv = F.conv2d(xt, wt, bias=shift, stride=2, padding=3)
    ref = F.relu(conv).permute(0, 2, 3, 1)
    check("sliding stem fwd", y[:, :H2, :W2], ref, 1e-2, 2e-2)
    # max-pool on the pitched tensor == max-pool on the dense copy, bit for bit
    oh, ow = (H2 + 2 - 3) // 2 + 1, (W2 + 2 - 3) // 2 + 1
    yd = y[:, :H2, :W2].contiguous()
    p1, a1 = torch.zeros(B, oh, ow, 64, dtype=BF, device="cuda"), torch.zeros(B, oh, ow, 64, dtype=torch.uint8, device="cuda")
    p2, a2 = torch.zeros_like(p1), torch.zeros_like(a1)
    ops.maxpool_fwd(y, p1, a1, B, H2, W2, 64, oh, ow, XH=HP, XW=WP)
    ops.maxpool_fwd(yd, p2, a2, B, H2, W2, 64, oh, ow)
    assert torch.equal(p1, p2) and torch.equal(a1, a2)
    dp = rnd(B, oh, ow, 64, seed=4).to(BF)
    dx1 = torch.full((B, HP, WP, 64), 3.0, dtype=BF, device="cuda")
    dx2 = torch.zeros(B, H2, W2, 64, dtype=BF, device="cuda")
    ops.maxpool_bwd(dp, a1, dx1, B, H2, W2, 64, oh, ow, XH=HP, XW=WP)
    ops.maxpool_bwd(dp, a2, dx2, B, H2, W2, 64, oh, ow)
    assert torch.equal(dx1[:, :H2, :W2], dx2)
    assert float(dx1[:, H2:].abs().max()) == 0.0 and float(dx1[:, :, W2:].abs().max()) == 0.0
    # weight gradient over the pitched gradient tensor (zeros in the wrapped rows / columns)
    dy = torch.zeros(B, HP, WP, 64, dtype=BF, device="cuda")
    dy[:, :H2, :W2] = rnd(B, H2, W2, 64, seed=5).to(BF)
    scale = rnd(64, seed=6).abs() + 0.5
    dW = torch.zeros(64, 256, dtype=F32, device="cuda")
    db = torch.zeros(64, dtype=F32, device="cuda")
    ops.wgrad(s2d, 16, dy, 64, M, 64, 256, g, dW, 256, rowscale=scale, dbias=db, a_kb_rows=WP, k_mask=True)
    torch.cuda.synchronize()
    gw, = torch.autograd.grad(conv, [wt], dy[:, :H2, :W2].float().permute(0, 3, 1, 2))
    ref16 = Engine._stem_to_s2d(None, (gw.permute(0, 2, 3, 1) * scale[:, None, None, None]).cpu()).reshape(64, 256).cuda()
    check("sliding stem wgrad", dW, ref16, 1e-2, 1e-2 * float(ref16.abs().max()))
    refb = dy.float().sum((0, 1, 2)) * scale
    check("sliding stem dbias", db, refb, 1e-2, 1e-2 * float(refb.abs().max()))
    # the columns that do not exist in the 7x7x3 kernel receive exactly zero
    exists = Engine._stem_to_s2d(None, torch.ones(1, 7, 7, 3)).reshape(256).bool().cuda()
    assert float(dW[:, ~exists].abs().max()) == 0.0


def _bits_to_bool(bits, N):
    """[M, N/8] uint8 -> [M, N] bool (bit n % 8 of byte n / 8)"""
    sh = torch.arange(8, device=bits.device, dtype=torch.uint8)
    return (((bits[:, :, None] >> sh[None, None, :]) & 1) > 0).reshape(bits.shape[0], N)


@pytest.mark.parametrize("M,N,K,bn", [(1000, 256, 64, 0), (4175, 64, 256, 0), (1300, 512, 128, 0), (777, 128, 512, 128), (20000, 256, 64, 256),
                                      (5000, 1024, 256, 0), (128 * 148 * 2 + 77, 64, 64, 64)])
def test_gemm_tc_one_bit_relu_masks(ops, M, N, K, bn):
    """detrb_igemm_t.out_bits / mask_bits (the backbone's ReLU masks at 1 bit per element): the forward epilogue writes
    bit = (result > 0) next to C; a data gradient that consumes the bits gives the bytes of the same launch with the bf16 mask."""
    if bn == 256 and ops.set_tc_persistent(2) != 2:
        ops.set_tc_persistent(0)
        pytest.skip("256-wide tiles exist only in the persistent kernel")
    A = rnd(M, K, seed=1).to(BF)
    W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
    bias, res = rnd(N, seed=3), rnd(M, N, seed=4).to(BF)
    C = torch.full((M + 2, N), 7.0, dtype=BF, device="cuda")
    bits = torch.full((M + 2, N // 8), 0xA5, dtype=torch.uint8, device="cuda")
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, residual=res, ldr=N, relu=True, C=C, ldc=N, out_bits=bits, ldob=N // 8,
              force_tc=bn)
    torch.cuda.synchronize()
    check("fwd", C[:M], F.relu(A.float() @ W.float().t() + bias + res.float()), 1e-2, 1e-2)
    assert torch.equal(_bits_to_bool(bits[:M], N), C[:M].float() > 0)
    assert int((bits[M:] != 0xA5).sum()) == 0 and float((C[M:].float() - 7.0).abs().max()) == 0       # guard rows untouched
    # consumer: same launch with the bf16 activation as mask and with its bits
    G = rnd(M, K, seed=6).to(BF)
    D1 = torch.empty(M, N, dtype=BF, device="cuda")
    D2 = torch.empty(M, N, dtype=BF, device="cuda")
    ops.igemm(G, W, M, N, K, K, K, ops.plain_geom(M, K), residual=res, ldr=N, mask=C, ldm=N, mask_scale=1.0, C=D1, ldc=N, force_tc=bn)
    ops.igemm(G, W, M, N, K, K, K, ops.plain_geom(M, K), residual=res, ldr=N, mask_bits=bits, ldmb=N // 8, mask_scale=1.0, C=D2, ldc=N,
              force_tc=bn)
    torch.cuda.synchronize()
    assert torch.equal(D1, D2)
    assert float(D1.float().abs().max()) > 0


def test_one_bit_masks_conv_and_direct_store_paths(ops):
    """out_bits from an im2col (3x3) forward, mask_bits in the strided scatter-accumulate data gradient of a 1x1/stride-2 shortcut
    (direct-store epilogue) and in a 3x3/stride-2 data gradient (parity-class sub-convolutions)"""
    B, H, Wd, Ci, Co = 2, 30, 44, 64, 128
    x = F.relu(rnd(B, H, Wd, Ci, seed=1)).to(BF)
    w = rnd(Co, 9 * Ci, scale=(9 * Ci) ** -0.5, seed=2).to(BF)
    M = B * H * Wd
    y = torch.empty(M, Co, dtype=BF, device="cuda")
    bits = torch.zeros(M, Co // 8, dtype=torch.uint8, device="cuda")
    g = dict(batch=B, IH=H, IW=Wd, Cin=Ci, OH=H, OW=Wd, KH=3, KW=3, stride=1, pad=1, mode=0)
    ops.igemm(x, w, M, Co, 9 * Ci, Ci, 9 * Ci, g, relu=True, C=y, ldc=Co, out_bits=bits, ldob=Co // 8)
    torch.cuda.synchronize()
    assert torch.equal(_bits_to_bool(bits, Co), y.float() > 0)
    # 1x1 / stride-2 shortcut data gradient: scatter to the even pixels of the input grid, accumulate, mask of the input activation
    xb = torch.zeros(M, Ci // 8, dtype=torch.uint8, device="cuda")
    xin = rnd(B, H, Wd, Ci, seed=3).to(BF)
    ops.igemm(xin, torch.eye(Ci, device="cuda").to(BF).contiguous(), M, Ci, Ci, Ci, Ci, ops.plain_geom(M, Ci), relu=True,
              C=torch.empty(M, Ci, dtype=BF, device="cuda"), ldc=Ci, out_bits=xb, ldob=Ci // 8)
    oh, ow = H // 2, Wd // 2
    Mo = B * oh * ow
    dy = rnd(Mo, Co, seed=4).to(BF)
    wd = rnd(Ci, Co, scale=Co ** -0.5, seed=5).to(BF)
    base = rnd(M, Ci, seed=6).to(BF)
    g1 = dict(batch=B, IH=oh, IW=ow, Cin=Co, OH=oh, OW=ow, KH=1, KW=1, stride=1, pad=0, mode=0)
    outs = []
    for kw in (dict(mask=F.relu(xin.float()).to(BF).view(M, Ci), ldm=Ci), dict(mask_bits=xb, ldmb=Ci // 8)):
        o = base.clone()
        ops.igemm(dy, wd, Mo, Ci, Co, Co, Co, g1, mask_scale=1.0, C=o, ldc=Ci, out_stride=2, SH=H, SW=Wd, accumulate=True, **kw)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and not torch.equal(outs[0], base)
    # 3x3 / stride-2 / pad-1 data gradient (four parity classes) with the bits of its input activation
    w3 = rnd(Ci, 9 * Co, scale=(9 * Co) ** -0.5, seed=7).to(BF)
    g3 = dict(batch=B, IH=oh, IW=ow, Cin=Co, OH=H, OW=Wd, KH=3, KW=3, stride=2, pad=1, mode=1)
    outs = []
    for kw in (dict(mask=F.relu(xin.float()).to(BF).view(M, Ci), ldm=Ci), dict(mask_bits=xb, ldmb=Ci // 8)):
        o = torch.zeros(M, Ci, dtype=BF, device="cuda")
        ops.igemm(dy, w3, M, Ci, 9 * Co, Co, 9 * Co, g3, mask_scale=1.0, C=o, ldc=Ci, **kw)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and float(outs[0].float().abs().max()) > 0


@pytest.mark.parametrize("M,N,K,epi", [(534400, 256, 64, "r"), (534400, 256, 64, "rb"), (534400, 64, 256, "b"), (534400, 64, 64, ""), (133600, 512, 128, "r"),
                                       (133600, 512, 128, "rb"), (534400, 128, 256, ""), (128 * 148 * 4 + 13, 256, 128, "rb"), (300, 128, 64, "r"),
                                       (128 * 5, 1024, 64, "rb"), (8400, 2048, 256, "d"), (8400, 1024, 128, "rd"), (33600, 1024, 256, "r")])
def test_gemm_stream_kernel(M, N, K, epi):
    """gemm_stream_kernel (weights resident in shared memory, in-place chunk slots) on the layer1 / layer2 1x1 shapes at full size,
    a ragged last tile, fewer tiles than CTAs, four column parts: against fp32 PyTorch on sampled rows, and bit for bit against the
    one-tile kernel on the whole output (same bf16 products, same fp32 accumulation order inside a k-block chain)."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import ops
    A = rnd(M, K, seed=1).to(BF)
    W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
    bias = rnd(N, seed=3)
    kw = {}
    if "r" in epi:
        kw.update(residual=rnd(M, N, seed=4).to(BF), ldr=N)
    if "b" in epi:
        kw.update(mask_bits=torch.randint(0, 256, (M, N // 8), dtype=torch.uint8, generator=torch.Generator().manual_seed(5)).cuda(), ldmb=N // 8,
                  mask_scale=1.0)
    if "d" in epi:                  # the transformer's epilogue: ReLU -> dropout (counter-based) -> + residual; no bit masks
        kw.update(drop_p=0.1, seed=11, site=3, seed_ptr=torch.tensor([77], dtype=torch.int64, device="cuda"))
    outs, bits = [], []
    for mode in (2, 0):
        olds, oldp = ops.set_tc_stream(mode), ops.set_tc_persistent(0)
        try:
            C = torch.full((M + 2, N), 7.0, dtype=BF, device="cuda")
            ob = torch.full((M + 2, N // 8), 0x5A, dtype=torch.uint8, device="cuda")
            okw = {} if "d" in epi else dict(out_bits=ob, ldob=N // 8)
            ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), bias=bias, relu=True, C=C, ldc=N, **okw, **kw)
            torch.cuda.synchronize()
        finally:
            ops.set_tc_stream(olds)
            ops.set_tc_persistent(oldp)
        outs.append(C)
        bits.append(ob)
    assert torch.equal(outs[0], outs[1]) and torch.equal(bits[0], bits[1])
    C = outs[0]
    assert float((C[M:].float() - 7.0).abs().max()) == 0 and int((bits[0][M:] != 0x5A).sum()) == 0
    rows = torch.cat([torch.arange(0, min(M, 300)), torch.randint(0, M, (2000,), generator=torch.Generator().manual_seed(6)),
                      torch.arange(max(0, M - 300), M)]).cuda()
    if "d" in epi:                  # dropout: exact agreement with the one-tile kernel above; statistics of the kept fraction here
        base = F.relu(A[rows].float() @ W.float().t() + bias)
        got = C[rows].float() - (kw["residual"][rows].float() if "r" in epi else 0)
        kept = (got.abs() > 1e-3) & (base > 0.05)
        frac = float(kept.sum()) / float((base > 0.05).sum())
        assert abs(frac - 0.9) < 0.01, frac
        return
    ref = A[rows].float() @ W.float().t() + bias
    if "r" in epi:
        ref = ref + kw["residual"][rows].float()
    ref = F.relu(ref)
    if "b" in epi:
        ref = torch.where(_bits_to_bool(kw["mask_bits"][rows], N), ref, torch.zeros_like(ref))
    check("stream", C[rows], ref, 1e-2, 1e-2)
    assert torch.equal(_bits_to_bool(bits[0][:M], N), C[:M].float() > 0)


@pytest.mark.parametrize("B,H,W", [(2, 20, 334), (1, 7, 100), (3, 33, 300), (1, 1, 5), (8, 200, 334), (2, 150, 129)])
def test_conv3x3_halo_kernel(B, H, W):
    """conv_halo.cu (3x3 / stride 1 / 64 -> 64 channels, every input row staged once, the nine taps as shifted descriptor views):
    forward (+ folded-BN shift, ReLU, 1-bit mask out) and data gradient (+ 1-bit mask in) against fp32 PyTorch, and bit for bit
    against the TMA-im2col kernel (same products, same accumulation order); 1..3 column strips, partial last strip, runs that start
    and end inside an image, fewer units than CTAs."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import ops
    C = 64
    x = rnd(B, H, W, C, seed=1).to(BF)
    w = rnd(C, 3, 3, C, scale=(9 * C) ** -0.5, seed=2).to(BF)
    shift = rnd(C, seed=3)
    M, K = B * H * W, 9 * C
    g = conv_geom(B, H, W, C, H, W, 3, 3, 1, 1)
    dy = rnd(B, H, W, C, seed=5).to(BF)
    wd = w.reshape(C, 9, C).permute(2, 1, 0).contiguous()
    mbits = torch.randint(0, 256, (M, C // 8), dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).cuda()
    gd = conv_geom(B, H, W, C, H, W, 3, 3, 1, 1, mode=1)
    res = {}
    for halo in (1, 0):
        old = ops.set_tc_halo(halo)
        try:
            y = torch.full((M + 1, C), 7.0, dtype=BF, device="cuda")
            ob = torch.full((M + 1, C // 8), 0x5A, dtype=torch.uint8, device="cuda")
            ops.igemm(x, w, M, C, K, C, K, g, bias=shift, relu=True, C=y, ldc=C, out_bits=ob, ldob=C // 8)
            dx = torch.full((M + 1, C), 7.0, dtype=BF, device="cuda")
            ops.igemm(dy, wd, M, C, K, C, K, gd, mask_bits=mbits, ldmb=C // 8, mask_scale=1.0, C=dx, ldc=C)
            torch.cuda.synchronize()
        finally:
            ops.set_tc_halo(old)
        res[halo] = (y, ob, dx)
    y, ob, dx = res[1]
    assert float((y[M:].float() - 7.0).abs().max()) == 0 and int((ob[M:] != 0x5A).sum()) == 0 and float((dx[M:].float() - 7.0).abs().max()) == 0
    ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias=shift, padding=1).permute(0, 2, 3, 1)).reshape(M, C)
    check("halo fwd", y[:M], ref, 1e-2, 2e-2)
    assert torch.equal(_bits_to_bool(ob[:M], C), y[:M].float() > 0)
    xt = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    o = F.conv2d(xt, w.float().permute(0, 3, 1, 2), padding=1)
    gx, = torch.autograd.grad(o, [xt], dy.float().permute(0, 3, 1, 2))
    refd = gx.permute(0, 2, 3, 1).reshape(M, C) * _bits_to_bool(mbits, C)
    check("halo dgrad", dx[:M], refd, 1e-2, 2e-2)
    for a, b in zip(res[1], res[0]):
        assert torch.equal(a, b)


@pytest.mark.parametrize("B,H,W,Ci,Co", [(2, 30, 44, 64, 128), (1, 25, 43, 128, 256), (8, 200, 334, 128, 128), (8, 100, 167, 256, 256)])
def test_stride2_scatter_through_the_dense_workspace(B, H, W, Ci, Co):
    """detrb_igemm_t.scratch: the two stride-2 scatter forms of the backbone's backward pass -- the data gradient of a 3x3 / stride-2
    convolution (four parity classes) and of a 1x1 / stride-2 shortcut (scatter-accumulate into the even pixels) -- written densely
    by the fast kernels and scattered by one coalesced pass, against the same launches without workspace (results scattered by the
    GEMM epilogue).  Identical arithmetic except one extra bf16 rounding of the shortcut term before it is added."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import ops
    M = B * H * W
    oh, ow = (H - 1) // 2 + 1, (W - 1) // 2 + 1
    Mo = B * oh * ow
    g = torch.Generator().manual_seed(9)
    dy = (torch.randn(Mo, Co, generator=g)).cuda().to(BF)
    bits = torch.randint(0, 256, (M, Ci // 8), dtype=torch.uint8, generator=g).cuda()
    scratch = torch.empty(M * Ci + 64, dtype=BF, device="cuda")
    # 3x3 / stride 2 / pad 1 data gradient
    w3 = (torch.randn(Ci, 9 * Co, generator=g) * (9 * Co) ** -0.5).cuda().to(BF)
    g3 = dict(batch=B, IH=oh, IW=ow, Cin=Co, OH=H, OW=W, KH=3, KW=3, stride=2, pad=1, mode=1)
    outs = []
    for sc in (None, scratch):
        o = torch.full((M + 1, Ci), 7.0, dtype=BF, device="cuda")
        ops.igemm(dy, w3, M, Ci, 9 * Co, Co, 9 * Co, g3, mask_bits=bits, ldmb=Ci // 8, mask_scale=1.0, C=o, ldc=Ci, scratch=sc)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]) and float(outs[0][:M].float().abs().max()) > 0 and float((outs[0][M:].float() - 7.0).abs().max()) == 0
    # 1x1 / stride-2 shortcut data gradient, accumulated into the even pixels
    wd = (torch.randn(Ci, Co, generator=g) * Co ** -0.5).cuda().to(BF)
    base = torch.randn(M + 1, Ci, generator=g).cuda().to(BF)
    g1 = dict(batch=B, IH=oh, IW=ow, Cin=Co, OH=oh, OW=ow, KH=1, KW=1, stride=1, pad=0, mode=0)
    outs = []
    for sc in (None, scratch):
        o = base.clone()
        ops.igemm(dy, wd, Mo, Ci, Co, Co, Co, g1, mask_bits=bits, ldmb=Ci // 8, mask_scale=1.0, C=o, ldc=Ci, out_stride=2, SH=H, SW=W, accumulate=True,
                  scratch=sc)
        outs.append(o)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][M:], base[M:]) and torch.equal(outs[1][M:], base[M:])
    odd = torch.ones(B, H, W, dtype=torch.bool, device="cuda")
    odd[:, ::2, ::2] = False
    assert torch.equal(outs[1][:M][odd.view(-1)], base[:M][odd.view(-1)])            # only the even pixels are touched
    check("shortcut scatter", outs[1][:M], outs[0][:M], 1e-2, 2e-2)
    assert not torch.equal(outs[1][:M], base[:M])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["plain256", "resid_bits256", "maskbits256", "dropout256", "plain128", "conv256", "conv256_dgrad", "conv128",
                                  "tail256"])
def test_gemm_pair_kernel(case):
    """gemm_pair_kernel (tcgen05 cta_group::2: one 256 x BN tile per CTA pair, each CTA stages its 128 rows of A and its half of W):
    bit for bit against the single-CTA persistent kernel (same products, same accumulation order, same epilogue), plus fp32 PyTorch.
    Odd counts of 128-row tiles (the peer CTA of the last pair entirely past M), residual ring, 1-bit masks in / out, dropout,
    im2col operands (3x3 convolution forward and data gradient), 128- and 256-wide tiles."""
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    from detr_tensorflow_b200 import ops
    kw, geom, ref = {}, None, None
    if case in ("plain256", "resid_bits256", "maskbits256", "dropout256", "plain128", "tail256"):
        M, N, K = {"plain256": (33600, 256, 1024), "resid_bits256": (33600, 512, 512), "maskbits256": (20000, 256, 2048),
                   "dropout256": (8400, 256, 2048), "plain128": (33600, 128, 1152), "tail256": (129, 256, 1024)}[case]
        A = rnd(M, K, seed=1).to(BF)
        W = rnd(N, K, scale=K ** -0.5, seed=2).to(BF)
        bias = rnd(N, seed=3)
        geom = ops.plain_geom(M, K)
        kw = dict(bias=bias, relu=case in ("plain256", "resid_bits256", "plain128", "tail256"))
        ref = A.float() @ W.float().t() + bias
        if case == "resid_bits256":
            R = rnd(M, N, seed=4).to(BF)
            kw.update(residual=R, ldr=N)
            ref = ref + R.float()
        if kw["relu"]:
            ref = F.relu(ref)
        if case == "maskbits256":
            mb = torch.randint(0, 256, (M, N // 8), dtype=torch.uint8, generator=torch.Generator().manual_seed(7)).cuda()
            kw.update(mask_bits=mb, ldmb=N // 8, mask_scale=1.25)
            ref = ref * _bits_to_bool(mb, N) * 1.25
        if case == "dropout256":
            R = rnd(M, N, seed=4).to(BF)
            kw.update(residual=R, ldr=N, drop_p=0.1, seed=1234, site=5)
            ref = None
        lda, ldw = K, K
    else:
        B, H, Wd, Cc = {"conv256": (8, 50, 84, 256), "conv256_dgrad": (3, 50, 84, 256), "conv128": (2, 100, 167, 128)}[case]
        mode = 1 if case == "conv256_dgrad" else 0
        x = rnd(B, H, Wd, Cc, seed=1).to(BF)
        w = rnd(Cc, 3, 3, Cc, scale=(9 * Cc) ** -0.5, seed=2).to(BF)
        M, N, K = B * H * Wd, Cc, 9 * Cc
        A, W, lda, ldw = x, w, Cc, K
        geom = conv_geom(B, H, Wd, Cc, H, Wd, 3, 3, 1, 1, mode=mode)
        if mode == 0:
            bias = rnd(Cc, seed=3)
            kw = dict(bias=bias, relu=True)
            ref = F.relu(F.conv2d(x.float().permute(0, 3, 1, 2), w.float().permute(0, 3, 1, 2), bias=bias, padding=1).permute(0, 2, 3, 1)).reshape(M, N)
        else:
            W = w.reshape(Cc, 9, Cc).permute(2, 1, 0).contiguous()
            xt = torch.zeros(B, Cc, H, Wd, device="cuda", requires_grad=True)
            o = F.conv2d(xt, w.float().permute(0, 3, 1, 2), padding=1)
            gx, = torch.autograd.grad(o, [xt], x.float().permute(0, 3, 1, 2))
            ref = gx.permute(0, 2, 3, 1).reshape(M, N)
    want_bits = case == "resid_bits256"
    res = {}
    oldp = ops.set_tc_persistent(2)
    try:
        for pair in (2, 0):
            old = ops.set_tc_pair(pair)
            try:
                y = torch.full((M + 1, N), 7.0, dtype=BF, device="cuda")
                ob = torch.full((M + 1, N // 8), 0x5A, dtype=torch.uint8, device="cuda")
                extra = dict(out_bits=ob, ldob=N // 8) if want_bits else {}
                ops.igemm(A, W, M, N, K, lda, ldw, geom, C=y, ldc=N, **kw, **extra)
                torch.cuda.synchronize()
            finally:
                ops.set_tc_pair(old)
            res[pair] = (y, ob)
    finally:
        ops.set_tc_persistent(oldp)
    y, ob = res[2]
    assert float((y[M:].float() - 7.0).abs().max()) == 0 and int((ob[M:] != 0x5A).sum()) == 0
    if ref is not None:
        check(case, y[:M], ref, 1e-2, 3e-2)
    if want_bits:
        assert torch.equal(_bits_to_bool(ob[:M], N), y[:M].float() > 0)
    assert torch.equal(res[2][0], res[0][0]) and torch.equal(res[2][1], res[0][1])
