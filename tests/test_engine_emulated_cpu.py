"""CPU tests of the host-side logic (engine orchestration, API mirror, optimizer glue, data-parallel path) with the
C ABI replaced by tests/cabi_emulator.py.  With fp32 storage the hand-written backward chain must reproduce the
oracle's autograd gradients to rounding error; with bf16 storage it must stay within bf16 noise on the forward."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import cabi_emulator
from oracle import detr_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NE, ND = 1, 2


@pytest.fixture()
def emu():
    cabi_emulator.install()
    cabi_emulator.set_act_dtype(torch.float32)
    yield cabi_emulator
    cabi_emulator.set_act_dtype(torch.bfloat16)
    cabi_emulator.uninstall()


def _setup(B=2, H=64, W=96, n=5, seed=1):
    P = O.init_params(seed=seed, num_encoder_layers=NE, num_decoder_layers=ND)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    tb, tc = O.synthetic_targets(B, n=n, seed=seed)
    return P, img, tb, tc


def rel(a, b):
    return float((a.float() - b.float()).norm() / (b.float().norm() + 1e-20))


def test_forward_loss_backward_match_oracle_fp32(emu):
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup()
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    out = model(img, training=False)
    with torch.no_grad():
        ref = O.detr_forward(P, img, num_encoder_layers=NE, num_decoder_layers=ND)
    assert out["pred_logits"].shape == (2, 100, 92) and out["pred_boxes"].shape == (2, 100, 4) and len(out["aux"]) == ND - 1
    assert rel(out["pred_logits"], ref["pred_logits"]) < 1e-4 and rel(out["pred_boxes"], ref["pred_boxes"]) < 1e-4
    eng = model.engine
    eng.set_targets(tb, tc)
    eng.zero_grads()
    eng.loss(91)
    total, log = eng.loss_dict()
    eng.backward()
    match = eng.a["match"].view(ND, 2, 100)
    _, ototal, olog, ograds = O.train_step(P, img, tb, tc, num_encoder_layers=NE, num_decoder_layers=ND, match_override=match)
    assert abs(float(total) - float(ototal)) < 1e-4 * abs(float(ototal))
    for k in olog:
        assert abs(float(log[k]) - float(olog[k])) < 1e-4 + 1e-4 * abs(float(olog[k])), k
    grads = eng.export_grads()
    assert set(grads) == set(ograds)
    for n_, g in grads.items():
        ref_g = ograds[n_]
        if float(ref_g.norm()) < 1e-6:
            assert float(g.norm()) < 1e-5, n_
        else:
            assert rel(g, ref_g) < 2e-3, (n_, rel(g, ref_g))
    # and the Hungarian assignment equals the oracle's (scipy) on the same fp32 outputs
    for l in range(ND):
        o_ = ref if l == ND - 1 else ref["aux"][l]
        for b in range(2):
            ti, pi, _, _, _, _ = O.hungarian_matching(tb[b], tc[b], o_["pred_boxes"][b], o_["pred_logits"][b])
            exp = -torch.ones(100, dtype=torch.int32)
            exp[pi] = ti.int()
            assert torch.equal(exp, match[l, b])


@pytest.mark.parametrize("case", ["a", "b", "ft"])
def test_engine_forward_vs_reference_code_golden(emu, case):
    """The product's host side (get_detr_model -> Engine orchestration: layouts, space-to-depth stem, sliding-window GEMM
    geometry, head wiring, fine-tuning heads) with the C ABI emulated in fp32, against activations produced by the REFERENCE'S
    OWN networks/*.py executed through get_detr_model() on the TensorFlow shim (tests/golden/make_golden_model.py) with the
    same seeded weights: equal to fp32 rounding.  Cases: batch 2; odd image sizes; include_top=False + nb_class=3."""
    import detr_tensorflow_b200 as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    seed, B, H, W, ne, nd, nbc = (int(v) for v in g[f"{case}_meta"])
    nbc = nbc or None
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd, nb_class=nbc)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=nbc is None, nb_class=nbc, params=P, dropout=0.0, num_encoder_layers=ne,
                             num_decoder_layers=nd, device="cpu")
    out = model(img, training=False)
    t = lambda k: torch.from_numpy(g[f"{case}_{k}"])
    assert rel(model.engine.feat.view(t("feat").shape), t("feat")) < 1e-5
    assert rel(out["pred_logits"], t("pred_logits")) < 1e-5 and rel(out["pred_boxes"], t("pred_boxes")) < 1e-5
    assert len(out["aux"]) == nd - 1
    for i, a in enumerate(out["aux"]):
        assert rel(a["pred_logits"], t(f"aux{i}_logits")) < 1e-5 and rel(a["pred_boxes"], t(f"aux{i}_boxes")) < 1e-5
    if nbc is not None:
        assert cfg.nlayers == ["cls_layer", "pos_layer"]            # config.add_nlayers, detr.py:103


def test_engine_resnet101_backbone_vs_reference_code_golden(emu):
    """get_detr_model(..., backbone='resnet101') (the BASELINE configs[3] backbone) on the fp32-emulated ABI against the
    reference's ResNet101Backbone (resnet_backbone.py:52-66) executed on the shim"""
    import detr_tensorflow_b200 as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    seed, B, H, W = (int(v) for v in g["r101_meta"])
    P = O.init_params(seed=seed, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1, device="cpu",
                             params=P, dropout=0.0)
    model(img, training=False)
    ref = torch.from_numpy(g["r101_feat"])
    assert rel(model.engine.feat.view(ref.shape), ref) < 1e-5


def test_engine_train_step_gradients_vs_reference_code_golden(emu):
    """The product's hand-written backward chain (Engine.loss + Engine.backward on the fp32-emulated C ABI) against the gradient
    of the REFERENCE'S OWN loss code through the REFERENCE'S OWN model code (make_golden_model.py::train_case): same Hungarian
    assignment, same total / per-term losses, and for every trainable variable the gradient norm, a seeded random projection
    and the full tensor of the small ones."""
    import detr_tensorflow_b200 as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    seed, B, H, W, ne, nd, n_t = (int(v) for v in g["train_meta"])
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    tb, tc = torch.from_numpy(g["train_t_bbox"]), torch.from_numpy(g["train_t_class"])
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=ne, num_decoder_layers=nd, device="cpu", params=P, dropout=0.0)
    model(img, training=False)
    eng = model.engine
    eng.set_targets(tb, tc)
    eng.zero_grads()
    eng.loss(91)
    total, log = eng.loss_dict()
    eng.backward()
    assert torch.equal(eng.a["match"].view(nd, B, 100).long(), torch.from_numpy(g["train_match"]))      # bit-exact assignment
    assert abs(float(total) - float(g["train_total"])) < 1e-4 * abs(float(g["train_total"]))
    for k, v in zip(g["train_loss_keys"].tolist(), g["train_loss_values"].tolist()):
        assert abs(float(log[k]) - v) < 1e-4 + 1e-4 * abs(v), (k, float(log[k]), v)
    grads = eng.export_grads()
    names = g["train_names"].tolist()
    assert set(names) - set(grads) == {"query_embed/kernel"} and set(grads) <= set(names)
    for i, n in enumerate(names):
        if n not in grads:
            continue
        gr = grads[n].float()
        ref_norm, ref_proj = float(g["train_grad_norms"][i]), float(g["train_grad_projs"][i])
        r = torch.randn(gr.shape, generator=torch.Generator().manual_seed(1000 + i))
        assert abs(float(gr.norm()) - ref_norm) <= 2e-3 * ref_norm + 1e-8, (n, float(gr.norm()), ref_norm)
        assert abs(float((gr * r).sum()) - ref_proj) <= 2e-3 * ref_norm * gr.numel() ** 0.5 + 1e-8, (n, float((gr * r).sum()), ref_proj)
        if "train_grad/" + n in g:
            full = torch.from_numpy(g["train_grad/" + n])
            assert float((gr - full).abs().max()) <= 2e-3 * float(full.abs().max()) + 1e-8, n


def test_training_config_defaults_vs_reference_code_golden():
    """TrainingConfig() attribute defaults, parser flags / defaults and add_nlayers against training_config.py executed as is
    (the learning rates are plain floats here, tf.Variable there; the lr flags are floats, type=bool in the reference)"""
    import json
    from detr_tensorflow_b200.training_config import TrainingConfig, training_config_parser
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    ref_attrs = json.loads(str(g["config_attrs_json"]))
    c = TrainingConfig()
    mine = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(c).items() if k != "data"}
    assert set(mine) == set(ref_attrs)
    for k, v in ref_attrs.items():
        assert mine[k] == pytest.approx(v) if isinstance(v, float) else mine[k] == v, (k, mine[k], v)
    ref_flags = json.loads(str(g["config_parser_json"]))
    flags = {a.dest: a.default for a in training_config_parser()._actions if a.dest != "help"}
    assert flags == ref_flags
    c.add_nlayers([type("L", (), {"name": "cls_layer"})(), type("L", (), {"name": "pos_layer"})()])
    assert c.nlayers == json.loads(str(g["config_nlayers_json"]))


def test_gradient_accumulation_cadence_vs_reference_code_golden(emu):
    """aggregate_grad_and_apply against the reference's own optimizers.py:137-163 executed with a recording optimizer
    (make_golden_model.py::accumulate_case): zero at step % k == 0, SUM of the micro-step gradients, apply at (step+1) % k == 0,
    nothing for a group whose train_<group> flag is off; k = 3 over 7 steps, and no accumulation."""
    import detr_tensorflow_b200 as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "model_golden.npz"))
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    for label, target_batch in (("acc3", 6), ("none", None)):
        cfg = D.TrainingConfig()
        cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 2, target_batch
        cfg.train_backbone, cfg.train_transformers = True, False
        model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
        opt = D.setup_optimizers(model, cfg)
        eng = model.engine
        lo, hi = eng.group_range["backbone"]
        calls = []
        eng.apply_group = lambda name, src, clip: calls.append((name, float(src[lo:lo + 3].sum()), float(src[lo + 3:lo + 7].sum())))
        steps = []
        for step in range(7):
            eng.grads.zero_()
            eng.grads[lo:lo + 3] = float(step + 1)
            eng.grads[lo + 3:lo + 7] = 10.0 * (step + 1)
            for name in ("backbone", "transformers"):
                n0 = len(calls)
                D.optimizers.aggregate_grad_and_apply(name, opt, None, step, cfg)
                if len(calls) > n0:
                    steps.append(step)
        assert all(c[0] == "backbone" for c in calls)
        assert steps == g[f"accum_{label}_steps"].tolist()
        np.testing.assert_allclose(np.array([c[1:] for c in calls]), g[f"accum_{label}_sums"], rtol=1e-6)


def test_bf16_storage_forward_noise_level(emu):
    import detr_tensorflow_b200 as D
    emu.set_act_dtype(torch.bfloat16)
    P, img, tb, tc = _setup()
    model = D.get_detr_model(D.TrainingConfig(), include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    out = model(img, training=False)
    with torch.no_grad():
        ref = O.detr_forward(P, img, num_encoder_layers=NE, num_decoder_layers=ND)
    assert rel(out["pred_logits"], ref["pred_logits"]) < 3e-2 and rel(out["pred_boxes"], ref["pred_boxes"]) < 2e-2


def test_training_fit_adam_and_accumulation(emu):
    """fit() over 2 micro-steps with target_batch = 2*batch_size: accumulate, then one Keras-Adam apply with
    per-variable clipnorm; compared with the oracle's restatement of optimizers.py."""
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    P2, img2, tb2, tc2 = _setup(B=1, H=32, W=48, n=4, seed=2)
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 1, 2
    cfg.train_backbone, cfg.train_transformers = True, True
    cfg.backbone_lr, cfg.transformers_lr = 1e-3, 1e-2
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
    opt = D.setup_optimizers(model, cfg)
    assert set(opt) >= {"backbone_optimizer", "transformers_optimizer", "nlayers_optimizer", "backbone_variables",
                        "transformers_variables", "nlayers_variables"}
    eng = model.engine
    # the body of fit() (training.py:41-54), keeping each micro-step's assignment so that the oracle can be evaluated
    # under the same matching (assignments of random-init predictions are unstable to 1e-6 perturbations)
    acc = None
    for step, (im, b_, c_) in enumerate(((img, tb, tc), (img2, tb2, tc2))):
        m_out, total_loss, log, gsteps = D.training.run_train_step(model, im, b_, c_, opt, cfg)
        assert set(gsteps) == {"backbone", "transformers", "nlayers"} and "backbone_lr" in log
        match = eng.a["match"].view(ND, 1, 100).clone()
        for name in gsteps:
            D.optimizers.aggregate_grad_and_apply(name, opt, gsteps[name]["gradients"], step, cfg)
        _, ototal, _, g = O.train_step(P, im, b_, c_, num_encoder_layers=NE, num_decoder_layers=ND, gradient_aggregate=2,
                                       match_override=match)
        assert abs(float(total_loss) - float(ototal)) < 1e-4 * abs(float(ototal))
        acc = g if acc is None else {k: acc[k] + g[k] for k in g}
    assert opt["backbone_optimizer"].iterations == 1 and opt["transformers_optimizer"].iterations == 1
    new = model.export_params()
    for name, g in acc.items():
        p = P[name].clone()
        lr = cfg.backbone_lr if O.param_group(name) == "backbone" else cfg.transformers_lr
        O.adam_clipnorm_step(p, g, torch.zeros_like(p), torch.zeros_like(p), 1, lr, 0.1)
        # Adam's first step moves every element by ~lr * sign(g): compare the update, not the value
        upd, ref_upd = new[name] - P[name], p - P[name]
        if float(g.norm()) < 1e-6:
            continue
        assert rel(upd, ref_upd) < 5e-2, (name, rel(upd, ref_upd))
    # frozen / ungrouped tensors untouched
    assert torch.equal(new["query_embed/kernel"], P["query_embed/kernel"])
    assert torch.equal(new["backbone/bn1/weight"], P["backbone/bn1/weight"])


def test_optimizer_step_split_around_the_stem_gradient(emu):
    """optimizer_step after backward(defer_tail=True) applies every variable but the stem kernel first and the stem kernel
    (the first chunks of the table) afterwards: identical to the one-call step"""
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    res = []
    for split in (False, True):
        model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
        eng = model.engine
        eng.forward(img, training=False)
        eng.set_targets(tb, tc)
        eng.set_lrs(1e-3, 1e-2)
        eng.set_enabled(True, True)
        for _ in range(2):
            eng.zero_grads()
            eng.loss(91)
            eng.backward()
            assert eng.stem_chunks == 2 and eng._tail is None
            if split:
                eng._tail = True                           # what backward(defer_tail=True) leaves behind on a GPU
            eng.optimizer_step(0.1)
            assert eng._tail is None
        res.append((eng.params.clone(), eng.adam_v.clone(), eng.steps.clone()))
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1]) and torch.equal(res[0][2], res[1][2])
    assert res[0][2][:2].tolist() == [2, 2]


def test_fit_on_step_hook_is_deferred_by_one_step_with_snapshot_values(emu, capsys):
    """fit() calls on_step(i) after step i+1 has been enqueued (a loss read-back then does not idle the GPU): every step is still
    reported exactly once, in order, with ITS OWN loss -- run_train_step returns snapshots, not views of the live loss buffers."""
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    _, img2, tb2, tc2 = _setup(B=1, H=32, W=48, n=4, seed=2)
    batches = [(img, tb, tc), (img2, tb2, tc2), (img, tb, tc)]

    def make():
        cfg = D.TrainingConfig()
        cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 1, None
        cfg.train_backbone, cfg.train_transformers = True, True
        cfg.backbone_lr, cfg.transformers_lr = 1e-3, 1e-2
        model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
        return cfg, model, D.setup_optimizers(model, cfg)
    # reference sequence: one step at a time, values read immediately
    cfg, model, opt = make()
    expect = []
    for step, (im, b_, c_) in enumerate(batches):
        _, total, log, gsteps = D.training.run_train_step(model, im, b_, c_, opt, cfg)
        expect.append((step, float(total), float(log["label_cost"]), float(log["l1_loss_0"])))
        held = (total, log["giou_loss"])
        for name in gsteps:
            D.optimizers.aggregate_grad_and_apply(name, opt, gsteps[name]["gradients"], step, cfg)
    # the values returned for the last step are snapshots: another step does not change them
    before = (float(held[0]), float(held[1]))
    D.training.run_train_step(model, img2, tb2, tc2, opt, cfg)
    assert (float(held[0]), float(held[1])) == before
    # fit: hook order, count and values
    cfg, model, opt = make()
    seen, enq = [], []
    orig = D.training.run_train_step

    def spy(*a, **k):
        enq.append(len(seen))                              # how many hooks had run when this step was enqueued
        return orig(*a, **k)
    D.training.run_train_step = spy
    try:
        D.training.fit(model, batches, opt, cfg, 0, None,
                       on_step=lambda s, t, l: seen.append((s, float(t), float(l["label_cost"]), float(l["l1_loss_0"]))))
    finally:
        D.training.run_train_step = orig
    assert enq == [0, 0, 1]                                # hook i runs after step i+1 was enqueued
    assert [x[0] for x in seen] == [0, 1, 2]
    for a, b in zip(seen, expect):
        assert a[0] == b[0] and all(abs(x - y) <= 1e-6 * abs(y) for x, y in zip(a[1:], b[1:])), (a, b)
    assert len({x[1] for x in seen}) == 3                  # three different losses (the optimizer moved, the batches differ)


def test_fit_loop_runs(emu, capsys):
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 1, 2
    cfg.train_backbone, cfg.train_transformers = True, True
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    opt = D.setup_optimizers(model, cfg)
    D.training.fit(model, [(img, tb, tc), (img, tb, tc)], opt, cfg, 0, None)
    D.training.eval(model, [(img, tb, tc)], cfg, None, evaluation_step=1)
    assert cfg.global_step == 2 and opt["backbone_optimizer"].iterations == 1
    outp = capsys.readouterr().out
    assert "Epoch: [0]" in outp and "Validation step: [0]" in outp


def test_group_gating(emu):
    import detr_tensorflow_b200 as D
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 1, None
    cfg.train_backbone, cfg.train_transformers = False, True
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    opt = D.setup_optimizers(model, cfg)
    D.training.fit(model, [(img, tb, tc)], opt, cfg, 0, None)
    new = model.export_params()
    assert torch.equal(new["backbone/layer1/0/conv1/kernel"], P["backbone/layer1/0/conv1/kernel"])      # not applied
    assert not torch.equal(new["class_embed/kernel"], P["class_embed/kernel"])
    assert opt["backbone_optimizer"].iterations == 0 and opt["transformers_optimizer"].iterations == 1


def test_standalone_losses_and_matching_api(emu, golden):
    import detr_tensorflow_b200 as D
    g = golden
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    logits, boxes = torch.from_numpy(g["l_logits"]), torch.from_numpy(g["l_boxes"])
    out = {"pred_logits": logits[5], "pred_boxes": boxes[5],
           "aux": [{"pred_logits": logits[i], "pred_boxes": boxes[i]} for i in range(5)]}
    total, losses = D.get_losses(out, g["l_t_bbox"], g["l_t_class"], cfg)
    assert abs(float(total) - float(g["l_total"])) < 2e-4 * float(g["l_total"])
    assert list(losses)[:6] == ["label_cost", "true_neg", "true_pos", "pos_accuracy", "giou_loss", "l1_loss"] and len(losses) == 36
    for k, v in zip(g["l_keys"], g["l_values"]):
        assert abs(float(losses[str(k)]) - float(v)) < 1e-4 + 1e-4 * abs(float(v)), k
    b = 0
    r = D.hungarian_matching(g["m_t_bbox"][b], g["m_t_class"][b], g["m_boxes"][b], g["m_logits"][b], device="cpu")
    assert np.array_equal(r[0].numpy(), g[f"m_t_indices_{b}"]) and np.array_equal(r[1].numpy(), g[f"m_p_indices_{b}"])
    assert np.array_equal(r[3].numpy(), g[f"m_p_selector_{b}"]) and bool(r[2].all())
    np.testing.assert_array_equal(r[4].numpy(), g[f"m_tb_{b}"])


def test_data_parallel_two_ranks_gloo(tmp_path):
    """world_size-2 gloo run of the N>1 path: each rank takes half of a global batch of 2, gradients are summed with
    one all-reduce and the loss uses GLOBAL normalisers -> must equal the single-process global-batch gradients."""
    script = os.path.join(ROOT, "tests", "dp_worker.py")
    port = 29500 + (os.getpid() % 2000)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    procs = [subprocess.Popen([sys.executable, script, str(r), "2", str(tmp_path)], env=env) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=600) == 0
    single = torch.load(os.path.join(tmp_path, "single.pt"))
    dp = torch.load(os.path.join(tmp_path, "dp_rank0.pt"))
    assert abs(dp["total_global"] - single["total"]) < 1e-4 * abs(single["total"])
    for k in single["grads"]:
        a, b = dp["grads"][k], single["grads"][k]
        if float(b.norm()) > 1e-6:
            assert rel(a, b) < 2e-3, (k, rel(a, b))
        # three overlapped bucket all-reduces (Engine.grad_buckets) == the one flat all-reduce, bit for bit
        assert torch.equal(dp["grads_bucketed"][k], a), k
        # ... and so does the public API (training.run_train_step under torch.distributed: direct staging, bucketed grads_step)
        assert torch.equal(dp["grads_api"][k], a), k
    assert abs(dp["total_api"] - single["total"]) < 1e-4 * abs(single["total"])      # logged loss = the GLOBAL loss on every rank


def test_finetune_heads_nlayers_group(emu):
    """get_detr_model(include_top=False, nb_class=N) (detr.py:94-114, the finetune_*.py scripts): Keras-Dense heads
    `cls_layer` / `pos_layer` in the 'nlayers' optimizer group (optimizers.py:39-43); only that group is applied when
    train_nlayers alone is set (finetune_voc.py:33-36 pattern), the other variables stay bit-identical."""
    import detr_tensorflow_b200 as D
    NB = 4                                                # 3 classes + background (class 0 in the VOC/CSV loaders)
    P = O.init_params(seed=3, num_encoder_layers=NE, num_decoder_layers=ND, nb_class=NB)
    assert P["cls_layer/kernel"].shape == (256, NB) and P["pos_layer/dense_2/kernel"].shape == (256, 4)
    assert "class_embed/kernel" not in P
    img = torch.randn(2, 64, 96, 3, generator=torch.Generator().manual_seed(3))
    tb, tc = O.synthetic_targets(2, n=4, num_classes=NB - 1, seed=3)
    tc = tc + (tb[:, :, 2:3] > 0).long()                  # class ids 1..NB-1, background = 0
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 0, 2, None
    cfg.train_backbone, cfg.train_transformers, cfg.train_nlayers = False, False, True
    cfg.nlayers_lr = 1e-2
    model = D.get_detr_model(cfg, include_top=False, nb_class=NB, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu",
                             params=P, dropout=0.0)
    assert cfg.nlayers == ["cls_layer", "pos_layer"]      # detr.py:103
    out = model(img, training=False)
    with torch.no_grad():
        ref = O.detr_forward(P, img, num_encoder_layers=NE, num_decoder_layers=ND)
    assert out["pred_logits"].shape == (2, 100, NB) and len(out["aux"]) == ND - 1
    assert rel(out["pred_logits"], ref["pred_logits"]) < 1e-4 and rel(out["pred_boxes"], ref["pred_boxes"]) < 1e-4
    opt = D.setup_optimizers(model, cfg)
    assert len(opt["nlayers_variables"]) == 8 and sum(v.numel() for v in opt["nlayers_variables"]) == 256 * NB + NB + 2 * (256 * 256 + 256) + 256 * 4 + 4
    eng = model.engine
    m_out, total_loss, log, gsteps = D.training.run_train_step(model, img, tb, tc, opt, cfg)
    match = eng.a["match"].view(ND, 2, 100).clone()
    _, ototal, olog, g = O.train_step(P, img, tb, tc, background_class=0, num_encoder_layers=NE, num_decoder_layers=ND,
                                      match_override=match)
    assert abs(float(total_loss) - float(ototal)) < 1e-4 * abs(float(ototal))
    grads = eng.export_grads()
    assert set(grads) == set(g)
    for n_ in g:
        if O.param_group(n_) == "nlayers":
            assert grads[n_].shape == P[n_].shape and rel(grads[n_], g[n_]) < 2e-3, n_
    for name in gsteps:
        D.optimizers.aggregate_grad_and_apply(name, opt, gsteps[name]["gradients"], 0, cfg)
    assert opt["nlayers_optimizer"].iterations == 1 and opt["backbone_optimizer"].iterations == 0
    new = model.export_params()
    for n_ in P:
        if O.param_group(n_) == "nlayers":
            p = P[n_].clone()
            O.adam_clipnorm_step(p, g[n_], torch.zeros_like(p), torch.zeros_like(p), 1, cfg.nlayers_lr, 0.1)
            assert rel(new[n_] - P[n_], p - P[n_]) < 5e-2, n_
        else:
            assert torch.equal(new[n_], P[n_]), n_         # frozen groups untouched
    # include_top=False without nb_class: the bare transformer output hs [L, B, 100, 256] (detr.py:177-179)
    bare = D.get_detr_model(D.TrainingConfig(), include_top=False, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu",
                            params=O.init_params(seed=3, num_encoder_layers=NE, num_decoder_layers=ND))
    assert tuple(bare(img, training=False).shape) == (ND, 2, 100, 256)


def test_inference_postprocess_and_uint8_input_host_logic(emu):
    """Host side of the N1/N2 rows with the C ABI emulated: get_model_inference (inference.py:68-95) and normalized_images
    (processing.py:6-23) against the vectors produced by the reference's own code; uint8 frames through the model equal the
    normalised float frames through the model; pad_labels reproduces the wire format."""
    import detr_tensorflow_b200 as D
    g = np.load(os.path.join(ROOT, "tests", "golden", "pipeline_golden.npz"))
    logits, boxes = torch.from_numpy(g["i_logits"]), torch.from_numpy(g["i_boxes"])
    for bg in (91, 0):
        for fmt in ("xy_center", "xyxy", "yxyx"):
            for b in range(3):
                pb, pl, ps = D.inference.get_model_inference({"pred_logits": logits[b:b + 1], "pred_boxes": boxes[b:b + 1]}, bg, fmt,
                                                              device="cpu")
                k = f"i_{bg}_{fmt}_{b}"
                assert np.array_equal(pl.numpy(), g[k + "_labels"])
                np.testing.assert_allclose(ps.numpy(), g[k + "_scores"], rtol=1e-6)
                np.testing.assert_allclose(pb.numpy(), g[k + "_bbox"], atol=1e-7)
    with pytest.raises(NotImplementedError):
        D.inference.get_model_inference({"pred_logits": logits[:1], "pred_boxes": boxes[:1]}, 91, "cxcy", device="cpu")
    cfg = D.TrainingConfig()
    for method in ("torch_resnet", "tf_resnet"):
        cfg.normalized_method = method
        out = D.data.normalized_images(g["n_img"], cfg, device="cpu")
        assert np.array_equal(out.numpy(), g[f"n_{method}"])                  # bit-exact (table built in float64 like the reference)
    for k in range(4):
        _, tb, tc = D.data.pad_labels(None, g[f"p_{k}_in_bbox"], g[f"p_{k}_in_class"])
        assert np.array_equal(tb, g[f"p_{k}_bbox"]) and np.array_equal(tc, g[f"p_{k}_class"]) and tc.dtype == np.int64
    with pytest.raises(ValueError):
        D.data.pad_labels(None, np.zeros((100, 4), np.float32), np.zeros((100, 1), np.int64))
    # uint8 frames straight into the model == normalised float frames into the model
    P, _, tb, tc = _setup(B=2, H=33, W=47)
    cfg.normalized_method = "torch_resnet"
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    u8 = torch.randint(0, 256, (2, 33, 47, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0))
    a = model(u8, training=False)["pred_logits"].clone()
    b = model(D.data.normalized_images(u8, cfg, device="cpu"), training=False)["pred_logits"].clone()
    assert torch.equal(a, b)


def test_detr_transform_host_logic(emu):
    """data/transformation.py on the emulated ABI.  The product's vectorised box transform equals the oracle's
    per-box restatement of transformation.py:11-34/117-142/178-186 for every geometry the sampler draws; the sampled maps keep
    the output inside the source unless zero_border is set; detr_transform keeps the reference's return order and dtypes."""
    import detr_tensorflow_b200 as D
    from oracle.resize_oracle import resize_affine_u8, transform_boxes as o_boxes
    cfg = D.TrainingConfig()
    cfg.image_size = (40, 56)
    rng = np.random.default_rng(3)
    g = np.random.default_rng(5)
    seen = set()
    for it in range(200):
        h, w = int(g.integers(20, 90)), int(g.integers(20, 90))
        fwd, zb = D.data.sample_geometry((h, w), cfg, True, rng)
        seen.add((fwd[0] < 0, zb))
        n = int(g.integers(0, 6))
        bb = np.concatenate([g.uniform(0.1, 0.9, (n, 2)), g.uniform(0.05, 0.6, (n, 2))], -1)
        cc = g.integers(1, 91, n)
        a, ac = D.data.transform_boxes(bb, cc, fwd, cfg.image_size, (h, w))
        b, bc = o_boxes(bb, cc, fwd, cfg.image_size, (h, w))
        assert np.array_equal(ac, bc) and a.shape == b.shape
        np.testing.assert_allclose(a, b, atol=1e-12)
        assert (a >= 0).all() and (a <= 1).all()
        if not zb:                              # the output frame maps into the source frame
            ax, bx, ay, by = D.data.transformation.inverse_map(fwd)
            xs = sorted([bx, ax * cfg.image_size[1] + bx])
            ys = sorted([by, ay * cfg.image_size[0] + by])
            assert xs[0] > -1e-6 and xs[1] < w + 1e-6 and ys[0] > -1e-6 and ys[1] < h + 1e-6
    assert len(seen) == 4                       # flips and zero-fill scalings both occur
    img = g.integers(0, 256, (30, 44, 3), dtype=np.uint8)
    out, nb, nc = D.data.detr_transform(img, np.array([[0.5, 0.5, 0.2, 0.2]]), np.array([7]), cfg, False, device="cpu")
    assert out.dtype == torch.float32 and tuple(out.shape) == (40, 56, 3) and nc.tolist() == [7]
    np.testing.assert_allclose(nb, [[0.5, 0.5, 0.2, 0.2]], atol=1e-12)          # plain resize keeps normalised boxes
    ref = resize_affine_u8([img], np.array([[44 / 56, 0, 30 / 40, 0]], np.float32), [0], 40, 56)[0]
    assert np.array_equal(out.numpy().astype(np.uint8), ref)


def test_checkpoint_roundtrip_and_torch_detr_name_mapping(emu, tmp_path):
    """save/load of parameters + Adam state (resume), and the original-DETR state_dict mapping: a torchvision
    resnet50 (the module the original DETR wraps as backbone.0.body) and nn.MultiheadAttention / nn.Linear / nn.LayerNorm
    modules with random weights, exported under the original key names, must drive the oracle to the torch modules' outputs."""
    import detr_tensorflow_b200 as D
    from detr_tensorflow_b200.networks import weights as Wt
    P, img, tb, tc = _setup(B=1, H=32, W=48, n=3)
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 1, None
    cfg.train_backbone, cfg.train_transformers = True, True
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
    opt = D.setup_optimizers(model, cfg)
    D.training.fit(model, [(img, tb, tc)] * 2, opt, cfg, 0, None)
    path = str(tmp_path / "ck.npz")
    Wt.save_checkpoint(model, path, cfg)
    cfg2 = D.TrainingConfig()
    cfg2.background_class, cfg2.batch_size, cfg2.target_batch = 91, 1, None
    cfg2.train_backbone, cfg2.train_transformers = True, True
    model2 = D.get_detr_model(cfg2, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", dropout=0.0,
                              weights=path, seed=123)
    Wt.load_checkpoint(model2, path, cfg2)
    e1, e2 = model.engine, model2.engine
    assert cfg2.global_step == 2 and torch.equal(e1.params, e2.params) and torch.equal(e1.adam_m, e2.adam_m)
    assert torch.equal(e1.adam_v, e2.adam_v) and torch.equal(e1.steps, e2.steps)
    # resumed training continues identically
    opt2 = D.setup_optimizers(model2, cfg2)
    D.training.fit(model, [(img, tb, tc)], opt, cfg, 0, None)
    D.training.fit(model2, [(img, tb, tc)], opt2, cfg2, 0, None)
    assert torch.equal(e1.params, e2.params)
    with pytest.raises(Exception):
        Wt.load_weights(model, "detr")                                   # the reference's bucket download: not offline

    # ---- original-DETR key names -> reference layouts
    import torchvision
    torch.manual_seed(0)
    r50 = torchvision.models.resnet50(weights=None).eval()
    for mod in r50.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.running_mean.normal_(0, 0.05)
            mod.running_var.uniform_(0.9, 1.1)
            mod.weight.data.normal_(1, 0.1)
            mod.bias.data.normal_(0, 0.05)
    sd = {"backbone.0.body." + k: v for k, v in r50.state_dict().items() if not k.startswith("fc.")}
    d = 256
    tr = torch.nn.ModuleDict({
        "input_proj": torch.nn.Conv2d(2048, d, 1), "query_embed": torch.nn.Embedding(100, d),
        "class_embed": torch.nn.Linear(d, 92)})
    for k, v in tr.state_dict().items():
        sd[k] = v
    for k in range(3):
        lin = torch.nn.Linear(d, d if k < 2 else 4)
        sd[f"bbox_embed.layers.{k}.weight"], sd[f"bbox_embed.layers.{k}.bias"] = lin.weight.data, lin.bias.data
    enc = torch.nn.TransformerEncoderLayer(d, 8, 2048, 0.0)             # same submodule names as the original DETR layer
    dec = torch.nn.TransformerDecoderLayer(d, 8, 2048, 0.0)
    for k, v in enc.state_dict().items():
        sd["transformer.encoder.layers.0." + k] = v
    for k, v in dec.state_dict().items():
        sd["transformer.decoder.layers.0." + k] = v
    fn = torch.nn.LayerNorm(d)
    fn.weight.data.normal_(1, 0.1)
    sd["transformer.decoder.norm.weight"], sd["transformer.decoder.norm.bias"] = fn.weight.data, fn.bias.data
    Pm = Wt.from_torch_detr_state_dict({"model": sd}, num_encoder_layers=1, num_decoder_layers=1)
    assert list(Pm) == list(O.param_shapes(num_encoder_layers=1, num_decoder_layers=1))
    x = torch.randn(1, 64, 96, 3)
    with torch.no_grad():
        feat = O.backbone_forward(Pm, x)
        t = x.permute(0, 3, 1, 2)
        t = r50.maxpool(r50.relu(r50.bn1(r50.conv1(t))))
        t = r50.layer4(r50.layer3(r50.layer2(r50.layer1(t))))
        assert rel(feat, t.permute(0, 2, 3, 1)) < 1e-5
        # encoder layer with pos = 0 equals torch's post-norm TransformerEncoderLayer; MHA / FFN / LN names verified
        src = torch.randn(2, 6, d)                                      # oracle layout [B, S, d]; torch layer is [S, B, d]
        ours = O.encoder_layer(Pm, "transformer/encoder/layer_0", src, torch.zeros(6, d))
        assert rel(ours, enc.eval()(src.transpose(0, 1)).transpose(0, 1)) < 1e-5
        mem, tgt = torch.randn(2, 9, d), torch.randn(2, 5, d)
        ours = O.decoder_layer(Pm, "transformer/decoder/layer_0", tgt, mem, torch.zeros(9, d), torch.zeros(5, d))
        assert rel(ours, dec.eval()(tgt.transpose(0, 1), mem.transpose(0, 1)).transpose(0, 1)) < 1e-5
        assert torch.equal(Pm["class_embed/kernel"], tr["class_embed"].weight) and Pm["input_proj/kernel"].shape == (1, 1, 2048, d)
    with pytest.raises(KeyError):
        Wt.from_torch_detr_state_dict({k: v for k, v in sd.items() if k != "query_embed.weight"}, num_encoder_layers=1,
                                      num_decoder_layers=1)


def test_parity_precision_pairs_forward_and_gradients():
    """precision="parity": bf16 storage, every activation / weight copy a PAIR of bf16 planes carved from one arena (the plane
    stride travels with every C-ABI call).  The engine's plumbing of the planes is checked here on the emulator (which stores
    real bf16 pairs): forward and per-variable gradients must sit at the 16-bit-mantissa level, two orders below plain bf16."""
    import detr_tensorflow_b200 as D
    cabi_emulator.install()
    cabi_emulator.set_act_dtype(torch.bfloat16)
    try:
        P, img, tb, tc = _setup()
        cfg = D.TrainingConfig()
        cfg.background_class = 91
        model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P,
                                 dropout=0.0, precision="parity")
        eng = model.engine
        out = model(img, training=False)
        assert eng.plane > 0 and eng.wplane > 0
        with torch.no_grad():
            ref = O.detr_forward(P, img, num_encoder_layers=NE, num_decoder_layers=ND)
            feat = O.backbone_forward(P, img)
        errs = (rel(eng.value(eng.feat).view(feat.shape), feat), rel(out["pred_logits"], ref["pred_logits"]), rel(out["pred_boxes"], ref["pred_boxes"]))
        print("parity-precision forward rel errors (feat, logits, boxes)", errs)
        assert max(errs) < 2e-4, errs
        eng.set_targets(tb, tc)
        eng.zero_grads()
        eng.loss(91)
        total, _ = eng.loss_dict()
        eng.backward()
        match = eng.a["match"].view(ND, 2, 100)
        _, ototal, _, ograds = O.train_step(P, img, tb, tc, num_encoder_layers=NE, num_decoder_layers=ND, match_override=match)
        assert abs(float(total) - float(ototal)) < 1e-4 * abs(float(ototal))
        grads = eng.export_grads()
        rels = sorted((rel(g, ograds[n_]), n_) for n_, g in grads.items() if float(ograds[n_].norm()) > 1e-6)
        worst, median = rels[-1], rels[len(rels) // 2]
        print("parity-precision gradient rel error: worst", worst, "median", median)
        # transformer / heads / layer4 sit at 1e-4; the early backbone reaches 5-7e-3: a ReLU whose input is within the 1e-5
        # forward error of zero flips its mask, and a fraction f of flipped elements costs sqrt(f) in relative L2 norm
        assert median[0] < 5e-4 and worst[0] < 1.5e-2, (worst, median)
    finally:
        cabi_emulator.uninstall()


def test_map_evaluation_vs_reference_code_golden(emu):
    """loss/compute_map.py (APDataObject / cal_map / calc_map mirrors, the batched MapEvaluator) on the emulated ABI against the
    reference's own compute_map.py (tests/golden/map_golden.npz): per (threshold, class) AP and the rounded summary, through
    both the per-image reference-style call (cal_map) and the batched wire-format path (one matching launch per batch)"""
    import detr_tensorflow_b200 as D
    from detr_tensorflow_b200.loss import compute_map as CM
    from oracle import map_oracle as MO
    g = np.load(os.path.join(ROOT, "tests", "golden", "map_golden.npz"))
    ncls, nimg = int(g["ncls"]), int(g["nimg"])
    names = [f"c{i}" for i in range(ncls)]
    thr = CM.IOU_THRESHOLDS
    ap = {"box": [[CM.APDataObject() for _ in names] for _ in thr], "mask": [[CM.APDataObject() for _ in names] for _ in thr]}
    for i in range(nimg):
        CM.cal_map(MO.yxyx_from_xcycwh(g[f"p_bbox_{i}"]), g[f"p_cls_{i}"], g[f"p_score_{i}"], None, MO.yxyx_from_xcycwh(g[f"t_bbox_{i}"]),
                   g[f"t_cls_{i}"], None, ap, thr)
    for a in range(len(thr)):
        for c in range(ncls):
            obj = ap["box"][a][c]
            assert obj.num_gt_positives == g["box_ngt"][a, c] and len(obj.data_points) == g["box_npts"][a, c]
            if g["box_ap"][a, c] >= 0:
                assert abs(obj.get_ap() - g["box_ap"][a, c]) < 1e-12, (a, c)
    maps = CM.calc_map(ap, thr, names)
    assert [str(k) for k in maps["box"].keys()] == g["box_map_keys"].tolist()
    np.testing.assert_allclose(list(maps["box"].values()), g["box_map_values"], atol=1e-9)
    np.testing.assert_allclose(list(maps["mask"].values()), g["mask_map_values"], atol=1e-9)
    # batched path: detections expressed as model outputs (one-hot-ish logits whose softmax maximum is the golden score is not
    # invertible exactly, so the batched matcher is fed post-processed tensors directly), targets in the padded wire format
    ev = CM.MapEvaluator(names)
    B, Q = nimg, 16
    boxes, labels, scores = torch.zeros(B, Q, 4), torch.zeros(B, Q, dtype=torch.int64), torch.zeros(B, Q)
    count = torch.zeros(B, dtype=torch.int32)
    tb, tc = torch.zeros(B, 100, 4), torch.zeros(B, 100, 1, dtype=torch.int64)
    for i in range(nimg):
        k, n = len(g[f"p_cls_{i}"]), len(g[f"t_cls_{i}"])
        boxes[i, :k] = torch.from_numpy(MO.yxyx_from_xcycwh(g[f"p_bbox_{i}"]))
        labels[i, :k], scores[i, :k], count[i] = torch.from_numpy(g[f"p_cls_{i}"]), torch.from_numpy(g[f"p_score_{i}"]), k
        tb[i, 0, 0] = n
        tb[i, 1:1 + n], tc[i, 1:1 + n, 0] = torch.from_numpy(g[f"t_bbox_{i}"]), torch.from_numpy(g[f"t_cls_{i}"])
    rank, tp, gtc = CM._match(boxes, labels, scores, count, tb, tc, None, True, thr, ncls)
    ev.batches.append((labels, scores, rank, tp, count, tb[:, 0, 0].clone(), tc[:, :, 0].clone()))
    maps2 = ev.summary()
    np.testing.assert_allclose(list(maps2["box"].values()), g["box_map_values"], atol=1e-9)
    assert gtc.tolist() == [int(sum((g[f"t_cls_{i}"] == c).sum() for i in range(nimg))) for c in range(ncls)]
