"""End-to-end GPU parity of the engine in the THROUGHPUT precision (plain bf16 activation storage, 8-bit mantissa) against the CPU
oracle, through the public API.  The fp32-tolerance parity tests are in test_parity_gpu.py (precision="parity").  Bounds here are
within 2x of the error levels measured on B200 (profiles/r02_parity_errors.txt): forward logits 1.1e-2 / boxes 4.4e-3 relative;
gradients are compared under the SAME assignment (match_override) because the Hungarian assignment of near-identical random-init
queries flips under bf16 perturbations; their error (median 6.7e-2, worst 0.38 on the early backbone) is ReLU-mask flips:
sqrt(forward error), see that file."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import detr_tensorflow_b200 as D
    return D


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


@pytest.fixture(scope="module")
def setup(D):
    from oracle import detr_oracle as O
    P = O.init_params(seed=1)
    img = torch.randn(2, 160, 224, 3, generator=torch.Generator().manual_seed(1))
    tb, tc = O.synthetic_targets(2, n=6, seed=1)
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0)
    return O, P, img, tb, tc, cfg, model


def test_forward_parity(D, setup):
    O, P, img, tb, tc, cfg, model = setup
    out = model(img, training=False)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.detr_forward(P, img)
    assert out["pred_logits"].shape == (2, 100, 92) and len(out["aux"]) == 5
    errs = {"logits": rel(out["pred_logits"], ref["pred_logits"]), "boxes": rel(out["pred_boxes"], ref["pred_boxes"]),
            "aux0": rel(out["aux"][0]["pred_logits"], ref["aux"][0]["pred_logits"])}
    print("forward rel errors", errs)
    # tolerance: bf16 storage through 50 conv layers + 12 transformer layers
    assert errs["logits"] < 2.5e-2 and errs["boxes"] < 1e-2 and errs["aux0"] < 2e-2, errs      # measured 1.1e-2 / 4.4e-3 / 8.5e-3
    # backbone feature map itself
    eng = model.engine
    with torch.no_grad():
        feat = O.backbone_forward(P, img)
    assert rel(eng.feat.view(feat.shape), feat) < 2e-2


def test_forward_vs_reference_code_golden(D):
    """The product's forward pass against tests/golden/model_golden.npz case "a": activations produced by the reference's own
    networks/*.py executed unmodified through get_detr_model() on the TensorFlow shim (tests/golden/make_golden_model.py), with
    the same seeded weights.  Tolerance: bf16 activation storage through 50 conv + 5 transformer layers (test_forward_parity's
    bounds with a little headroom: the 3x4 feature map averages over 12 tokens only)."""
    import os
    import numpy as np
    from oracle import detr_oracle as O
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz"))
    seed, B, H, W, ne, nd, _ = (int(v) for v in g["a_meta"])
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=ne, num_decoder_layers=nd)
    out = model(img, training=False)
    torch.cuda.synchronize()
    ref = {k: torch.from_numpy(g[f"a_{k}"]) for k in ("feat", "pred_logits", "pred_boxes", "aux0_logits", "aux1_boxes")}
    assert len(out["aux"]) == nd - 1
    errs = {"feat": rel(model.engine.feat.view(ref["feat"].shape), ref["feat"]), "logits": rel(out["pred_logits"], ref["pred_logits"]),
            "boxes": rel(out["pred_boxes"], ref["pred_boxes"]), "aux0_logits": rel(out["aux"][0]["pred_logits"], ref["aux0_logits"]),
            "aux1_boxes": rel(out["aux"][1]["pred_boxes"], ref["aux1_boxes"])}
    print("rel errors vs reference-code golden", errs)
    # measured: feat 7.7e-3, logits 1.4e-2, boxes 4.1e-3, aux0 logits 1.2e-2, aux1 boxes 3.9e-3
    assert errs["feat"] < 1.6e-2 and errs["logits"] < 3e-2 and errs["boxes"] < 1e-2 and errs["aux0_logits"] < 2.5e-2 and errs["aux1_boxes"] < 1e-2, errs


def test_baseline_config_c1_forward_480x640(D):
    """BASELINE.json configs[0] (1 synthetic 480x640 image, forward only) through get_detr_model()/model(): same output dict as
    the CPU oracle run of the same configuration, within the bf16-storage tolerance of test_forward_parity"""
    from oracle import detr_oracle as O
    P = O.init_params(seed=0)
    img = torch.randn(1, 480, 640, 3, generator=torch.Generator().manual_seed(0))
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0)
    out = model(img, training=False)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.detr_forward(P, img)
    assert out["pred_logits"].shape == (1, 100, 92) and out["pred_boxes"].shape == (1, 100, 4) and len(out["aux"]) == 5
    assert (model.engine.fh, model.engine.fw) == (15, 20)
    assert rel(out["pred_logits"], ref["pred_logits"]) < 3e-2 and rel(out["pred_boxes"], ref["pred_boxes"]) < 1.2e-2


def test_matcher_exact_on_engine_outputs_and_loss_parity(D, setup):
    O, P, img, tb, tc, cfg, model = setup
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    eng = model.engine
    out = model(img, training=False)
    eng.set_targets(tb, tc)
    cost = eng.match(want_cost=True)
    eng.loss(91, with_grad=False)
    total, log = eng.loss_dict()
    torch.cuda.synchronize()
    L, B, Q = eng.ndec, 2, 100
    assert int(eng.a["status"].abs().sum()) == 0
    pi, ti = eng.a["p_indices"].cpu().view(L, B, Q), eng.a["t_indices"].cpu().view(L, B, Q)
    cost = cost.cpu().view(L, B, Q, 100)
    for l in range(L):
        for b in range(B):
            n = int(tb[b, 0, 0])
            rows, cols = linear_sum_assignment(cost[l, b, :, :n].numpy())
            assert np.array_equal(pi[l, b, :n].numpy(), rows) and np.array_equal(ti[l, b, :n].numpy(), cols)   # bit-exact
    # loss values: oracle evaluated on the engine's own outputs with the engine's assignment
    o = {k: (v.float().cpu() if torch.is_tensor(v) else [{kk: vv.float().cpu() for kk, vv in a.items()} for a in v]) for k, v in out.items()}
    ototal, olog = O.get_losses(o, tb, tc, 91, eng.a["match"].cpu().view(L, B, Q))
    assert abs(float(total) - float(ototal)) < 1e-3 * abs(float(ototal))
    for k in olog:
        assert abs(float(log[k]) - float(olog[k])) < 1e-3 + 1e-3 * abs(float(olog[k])), k


def test_gradient_parity_under_same_assignment(D, setup):
    O, P, img, tb, tc, cfg, model = setup
    eng = model.engine
    model(img, training=False)
    eng.set_targets(tb, tc)
    eng.zero_grads()
    eng.loss(91)
    eng.backward()
    torch.cuda.synchronize()
    match = eng.a["match"].cpu().view(eng.ndec, 2, 100)
    total, _ = eng.loss_dict()
    _, ototal, _, og = O.train_step(P, img, tb, tc, match_override=match)
    assert abs(float(total) - float(ototal)) < 3e-2 * abs(float(ototal))
    g = eng.export_grads()
    rels = {n: rel(g[n], og[n]) for n in g if float(og[n].norm()) > 1e-6}
    worst = sorted(rels.items(), key=lambda kv: -kv[1])[:8]
    print("worst gradient rel errors", worst)
    heads = [n for n in rels if n.startswith(("class_embed", "bbox_embed"))]
    assert max(rels[n] for n in heads) < 0.15, [(n, rels[n]) for n in heads]
    # ReLU-mask flips: relative L2 error = sqrt(fraction of flipped elements) ~ sqrt(forward error 1e-2); measured median 6.7e-2,
    # worst 0.38 (early backbone).  The fp32-class check of the same chain is test_parity_gpu.py::test_parity_gradients_vs_oracle_full_model
    import numpy as np
    print("gradient rel errors: median", float(np.median(list(rels.values()))), "max", max(rels.values()))
    assert float(np.median(list(rels.values()))) < 0.14 and max(rels.values()) < 0.75, worst


def test_training_mode_steps_graph_and_dropout(D, setup):
    O, P, img, tb, tc, cfg, _ = setup
    cfg2 = D.TrainingConfig()
    cfg2.background_class = 91
    model = D.get_detr_model(cfg2, include_top=True, params=P, dropout=0.1)
    eng = model.engine
    a = model(img, training=True)["pred_logits"].clone()
    eng.seed_dev.add_(1)
    b = model(img, training=True)["pred_logits"].clone()
    c = model(img, training=False)["pred_logits"].clone()
    assert float((a - b).abs().max()) > 1e-4          # different dropout masks
    assert rel(a, c) < 0.5
    eng.set_targets(tb, tc)
    eng.set_lrs(1e-5, 1e-4)
    eng.set_enabled(True, True)
    before = eng.params.clone()
    replay = eng.capture_train_step(91, 0.1)
    losses = []
    for _ in range(3):
        replay()
        losses.append(float(eng.a["total"][0]))
    torch.cuda.synchronize()
    assert all(l == l and l < 1e4 for l in losses), losses
    assert int(eng.steps[0]) >= 3 and float((eng.params - before).abs().max()) > 0
    assert eng.launches_per_step > 100


def test_multi_tensor_weight_refresh(D, setup):
    """the single-launch refresh (prep_weights_multi) must produce exactly the per-layer layouts"""
    O, P, img, tb, tc, cfg, model = setup
    eng = model.engine
    eng.refresh_weights()
    torch.cuda.synchronize()
    for name in ("backbone/conv1", "backbone/layer1/0/conv2", "backbone/layer3/0/downsample", "input_proj", "class_embed",
                 "bbox_embed_2", "transformer/encoder/layer_0/self_attn/in_proj", "transformer/decoder/layer_5/linear2"):
        s = eng.slots[name]
        w = s.master.view(s.N, s.taps, s.Cin)
        if s.fold is not None:
            w = w * s.fold[:, None, None]
        wb = w.to(torch.bfloat16)
        assert torch.equal(s.Wf.view(s.N, s.taps, s.Cin), wb), name
        if s.Wd is not None:
            assert torch.equal(s.Wd[:, :, :s.N], wb.permute(2, 1, 0)), name
            assert float(s.Wd[:, :, s.N:].float().abs().max()) == 0 if s.ldd > s.N else True


def test_finetune_heads_nlayers_gpu(D):
    """detr.py:94-114 fine-tuning model (nb_class heads, Keras Dense [in,out] kernels, 'nlayers' group) on the CUDA path:
    forward parity, gradients of the new layers under the same assignment, and an nlayers-only Adam apply."""
    from oracle import detr_oracle as O
    NB = 4
    P = O.init_params(seed=5, nb_class=NB)
    img = torch.randn(2, 128, 160, 3, generator=torch.Generator().manual_seed(5))
    tb, tc = O.synthetic_targets(2, n=4, num_classes=NB - 1, seed=5)
    tc = tc + (tb[:, :, 2:3] > 0).long()
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 0, 2, None
    cfg.train_backbone, cfg.train_transformers, cfg.train_nlayers = False, False, True
    cfg.nlayers_lr = 1e-2
    model = D.get_detr_model(cfg, include_top=False, nb_class=NB, params=P, dropout=0.0)
    assert cfg.nlayers == ["cls_layer", "pos_layer"]
    out = model(img, training=False)
    with torch.no_grad():
        ref = O.detr_forward(P, img)
    assert out["pred_logits"].shape == (2, 100, NB) and len(out["aux"]) == 5
    # tolerance: bf16 storage through the whole network (same as test_forward_parity)
    print("bf16 forward rel errors (logits, boxes)", rel(out["pred_logits"], ref["pred_logits"]), rel(out["pred_boxes"], ref["pred_boxes"]))
    assert rel(out["pred_logits"], ref["pred_logits"]) < 4e-2 and rel(out["pred_boxes"], ref["pred_boxes"]) < 2e-2
    opt = D.setup_optimizers(model, cfg)
    eng = model.engine
    before = model.export_params()
    m_out, total_loss, log, gsteps = D.training.run_train_step(model, img, tb, tc, opt, cfg)
    match = eng.a["match"].view(6, 2, 100).clone().cpu()
    _, ototal, olog, g = O.train_step(P, img, tb, tc, background_class=0, match_override=match)
    assert abs(float(total_loss) - float(ototal)) < 3e-2 * abs(float(ototal))
    grads = eng.export_grads()
    for n_ in g:
        if O.param_group(n_) == "nlayers":
            assert rel(grads[n_], g[n_]) < 0.15, (n_, rel(grads[n_], g[n_]))      # head gradients: bf16 noise level
    for name in gsteps:
        D.optimizers.aggregate_grad_and_apply(name, opt, gsteps[name]["gradients"], 0, cfg)
    torch.cuda.synchronize()
    assert opt["nlayers_optimizer"].iterations == 1 and opt["backbone_optimizer"].iterations == 0
    after = model.export_params()
    for n_ in P:
        if O.param_group(n_) == "nlayers":
            assert float((after[n_] - before[n_]).abs().max()) > 0, n_
        else:
            assert torch.equal(after[n_], before[n_]), n_


def test_resnet101_backbone_forward(D):
    """BASELINE configs[3]: DETR-R101 (23 bottlenecks in layer3, resnet_backbone.py:52-66) through the same kernels."""
    from oracle import detr_oracle as O
    P = O.init_params(seed=2, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1)
    img = torch.randn(1, 128, 160, 3, generator=torch.Generator().manual_seed(2))
    model = D.get_detr_model(D.TrainingConfig(), include_top=True, params=P, dropout=0.0, backbone="resnet101",
                             num_encoder_layers=1, num_decoder_layers=1)
    out = model(img, training=False)
    with torch.no_grad():
        ref = O.detr_forward(P, img, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1)
        feat = O.backbone_forward(P, img, "resnet101")
    eng = model.engine
    assert len(eng.blocks) == 33 and eng.total > 45_000_000          # 17 extra layer3 blocks (x 1 114 112 parameters)
    assert rel(eng.feat.view(feat.shape), feat) < 4e-2
    print("bf16 forward rel errors (logits, boxes)", rel(out["pred_logits"], ref["pred_logits"]), rel(out["pred_boxes"], ref["pred_boxes"]))
    assert rel(out["pred_logits"], ref["pred_logits"]) < 4e-2 and rel(out["pred_boxes"], ref["pred_boxes"]) < 2e-2
    # one full train step runs (backward through 33 blocks, optimizer over the larger arena)
    tb, tc = O.synthetic_targets(1, n=3, seed=2)
    eng.set_targets(tb, tc)
    eng.set_lrs(1e-5, 1e-4)
    eng.set_enabled(True, True)
    eng.train_step(91, 0.1)
    torch.cuda.synchronize()
    t = float(eng.a["total"][0])
    assert t == t and 0 < t < 1e3


def test_streaming_halo_and_bit_mask_kernels_inside_the_full_size_step(D):
    """Round-2 kernels in situ: one 800x1333 image (BASELINE configs[1] resolution, every layer at its benchmark row count per
    image) through forward + loss + backward with the streaming GEMM / halo conv kernels ON (wherever supported) and OFF (one-tile /
    persistent / im2col kernels, same 1-bit masks).  The kernels compute the same bf16 products in the same order, so the forward
    is identical bit for bit; the weight gradients differ only by the order of their fp32 atomics."""
    from oracle import detr_oracle as O
    from detr_tensorflow_b200 import ops
    P = O.init_params(seed=2)
    img = torch.randn(1, 800, 1333, 3, generator=torch.Generator().manual_seed(2))
    tb, tc = O.synthetic_targets(1, n=7, seed=2)
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    res = {}
    for on in (1, 0):
        olds, oldh = ops.set_tc_stream(2 * on), ops.set_tc_halo(on)          # 2: the streaming kernel wherever it is supported
        try:
            model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0)
            eng = model.engine
            out = model(img, training=False)
            logits, boxes = out["pred_logits"].clone(), out["pred_boxes"].clone()
            eng.set_targets(tb, tc)
            eng.zero_grads()
            eng.loss(91)
            eng.backward()
            torch.cuda.synchronize()
            res[on] = (logits, boxes, float(eng.loss_dict()[0]), {k: v.clone() for k, v in eng.export_grads().items()})
        finally:
            ops.set_tc_stream(olds)
            ops.set_tc_halo(oldh)
        del model, eng
        torch.cuda.empty_cache()
    assert torch.equal(res[1][0], res[0][0]) and torch.equal(res[1][1], res[0][1])
    assert abs(res[1][2] - res[0][2]) <= 1e-6 * abs(res[0][2])
    worst = max((rel(res[1][3][n], res[0][3][n]), n) for n in res[0][3] if float(res[0][3][n].norm()) > 1e-8)
    print("new kernels on vs off: worst gradient difference", worst)
    assert worst[0] < 1e-3, worst
