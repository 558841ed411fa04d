"""The public API rows that round 1 only exercised on the CPU emulator, now through the CUDA path: gradient accumulation
(optimizers.py:137-163, `target_batch`), eval / run_val_step (training.py:28-32, 68-87), the standalone get_losses /
hungarian_matching wrappers (loss.py:22, hungarian_matching.py:163), checkpoint save / resume, the matcher-status poison."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NE, ND = 1, 2


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import detr_tensorflow_b200 as D
    return D


def rel(a, b):
    a, b = a.float().cpu(), b.float().cpu()
    return float((a - b).norm() / (b.norm() + 1e-20))


def _cfg(D, batch, target_batch):
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, batch, target_batch
    cfg.train_backbone, cfg.train_transformers = True, True
    return cfg


def test_accumulate_kernel(D):
    from detr_tensorflow_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    n = 4 * 100003
    acc, a, b = torch.full((n,), 7.0, device="cuda"), torch.randn(n, device="cuda", generator=g), torch.randn(n, device="cuda", generator=g)
    ops.accumulate(acc, a, n, True)
    ops.accumulate(acc, b, n, False)
    torch.cuda.synchronize()
    assert torch.equal(acc, a + b)


def test_gradient_accumulation_target_batch_on_device(D):
    """target_batch = 2 * batch_size (optimizers.py:137-163): two micro-steps accumulate on device (detrb_accumulate), the
    optimizer applies once, on the SUM of the two micro-step gradients each computed from total_loss / 2 (training.py:20);
    parity precision + dropout 0 so that the applied update can be checked against the oracle's Adam on the summed gradients."""
    from oracle import detr_oracle as O
    P = O.init_params(seed=3, num_encoder_layers=NE, num_decoder_layers=ND)
    imgs = [torch.randn(1, 96, 128, 3, generator=torch.Generator().manual_seed(30 + i)) for i in range(2)]
    tgts = [O.synthetic_targets(1, n=4, seed=30 + i) for i in range(2)]
    cfg = _cfg(D, 1, 2)
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=NE, num_decoder_layers=ND, precision="parity")
    opt = D.setup_optimizers(model, cfg)
    eng = model.engine
    p0 = eng.params.clone()
    sums = None
    for step in range(2):
        m_out, total, log, gsteps = D.training.run_train_step(model, imgs[step], tgts[step][0], tgts[step][1], opt, cfg)
        gnow = eng.grads.clone()
        sums = gnow if sums is None else sums + gnow
        for name in gsteps:
            D.optimizers.aggregate_grad_and_apply(name, opt, gsteps[name]["gradients"], step, cfg)
        torch.cuda.synchronize()
        if step == 0:
            assert torch.equal(eng.params, p0), "no apply before the accumulation window is full"
            assert torch.equal(eng.acc, gnow)
    assert torch.equal(eng.acc, sums)                                     # SUM of the micro-step gradients, on device
    assert opt["backbone_optimizer"].iterations == 1 and opt["transformers_optimizer"].iterations == 1
    assert float((eng.params - p0).abs().max()) > 0
    # the oracle: gradients of total / 2 per micro-batch (same assignment), summed, clipped per variable, one Keras-Adam step
    og = None
    for step in range(2):
        _, _, _, g_ = O.train_step(P, imgs[step], tgts[step][0], tgts[step][1], num_encoder_layers=NE, num_decoder_layers=ND,
                                   gradient_aggregate=2)
        og = g_ if og is None else {k: (og[k] + g_[k] if g_[k] is not None else og[k]) for k in og}
    after = model.export_params()
    worst_g, worst_u = 0.0, 0.0
    for n_, gsum in og.items():
        grp = O.param_group(n_)
        if gsum is None or grp is None or float(gsum.norm()) < 1e-6:
            continue
        acc = eng._to_ref_layout(n_, eng.acc)                             # the accumulated gradient the optimizer was given
        worst_g = max(worst_g, rel(acc, gsum))
        lr = cfg.backbone_lr if grp == "backbone" else cfg.transformers_lr
        p_ref = P[n_].clone()
        O.adam_clipnorm_step(p_ref, gsum, torch.zeros_like(p_ref), torch.zeros_like(p_ref), 1, lr, cfg.gradient_norm_clipping)
        # first Adam step: update = -lr * g / (|g| + eps'), a sign-like quantity: compare where the gradient is significant
        sig = gsum.abs() > 0.2 * gsum.abs().max()
        upd_ref, upd = (p_ref - P[n_])[sig], (after[n_] - P[n_])[sig]
        worst_u = max(worst_u, float((upd - upd_ref).abs().max()) / lr)
    print("accumulated gradient vs oracle: worst rel", worst_g, "; Adam update vs oracle: worst |diff| / lr", worst_u)
    assert worst_g < 3e-2 and worst_u < 5e-2


def test_eval_and_run_val_step_on_device(D, capsys):
    """training.eval / run_val_step (training.py:28-32, 68-87): forward(training=False) + losses, no gradients, no update"""
    from oracle import detr_oracle as O
    P = O.init_params(seed=2, num_encoder_layers=NE, num_decoder_layers=ND)
    img = torch.randn(2, 96, 128, 3, generator=torch.Generator().manual_seed(2))
    tb, tc = O.synthetic_targets(2, n=5, seed=2)
    cfg = _cfg(D, 2, None)
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.1, num_encoder_layers=NE, num_decoder_layers=ND, precision="parity")
    eng = model.engine
    p0 = eng.params.clone()
    m_out, total, log = D.training.run_val_step(model, img, tb, tc, cfg)
    torch.cuda.synchronize()
    with torch.no_grad():
        ref = O.detr_forward(P, img, num_encoder_layers=NE, num_decoder_layers=ND)     # eval mode: dropout off
    ototal, olog = O.get_losses(ref, tb, tc, 91)
    assert rel(m_out["pred_logits"], ref["pred_logits"]) < 5e-4
    assert abs(float(total) - float(ototal)) < 1e-3 * abs(float(ototal))
    for k in olog:
        assert abs(float(log[k]) - float(olog[k])) < 1e-3 + 1e-3 * abs(float(olog[k])), k
    D.training.eval(model, [(img, tb, tc)] * 3, cfg, None, evaluation_step=2)
    text = capsys.readouterr().out
    assert "Validation step: [0]" in text and torch.equal(eng.params, p0)


def test_standalone_loss_and_matching_wrappers_on_device(D):
    """D.get_losses / D.hungarian_matching (the functions a user of the reference calls directly) against the reference-code
    golden vectors (tests/golden/loss_golden.npz: loss.py / hungarian_matching.py / bbox.py executed with the real scipy)"""
    g = np.load(os.path.join(ROOT, "tests", "golden", "loss_golden.npz"))
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    logits, boxes = torch.from_numpy(g["l_logits"]), torch.from_numpy(g["l_boxes"])          # [6,B,Q,C], [6,B,Q,4]
    out = {"pred_logits": logits[5], "pred_boxes": boxes[5], "aux": [{"pred_logits": logits[i], "pred_boxes": boxes[i]} for i in range(5)]}
    total, losses = D.get_losses(out, g["l_t_bbox"], g["l_t_class"], cfg)
    assert abs(float(total) - float(g["l_total"])) < 2e-4 * float(g["l_total"])
    assert list(losses)[:6] == ["label_cost", "true_neg", "true_pos", "pos_accuracy", "giou_loss", "l1_loss"] and len(losses) == 36
    for k, v in zip(g["l_keys"], g["l_values"]):
        assert abs(float(losses[str(k)]) - float(v)) < 1e-4 + 1e-4 * abs(float(v)), k
    for b in range(6):                       # hungarian_matching on one image at a time, the reference's 6-tuple
        r = D.hungarian_matching(g["m_t_bbox"][b], g["m_t_class"][b], g["m_boxes"][b], g["m_logits"][b])
        assert np.array_equal(r[0].cpu().numpy(), g[f"m_t_indices_{b}"]) and np.array_equal(r[1].cpu().numpy(), g[f"m_p_indices_{b}"])
        assert np.array_equal(r[3].cpu().numpy(), g[f"m_p_selector_{b}"]) and bool(r[2].all())
        np.testing.assert_array_equal(r[4].cpu().numpy(), g[f"m_tb_{b}"])


def test_hungarian_matching_wrapper_and_nan_status(D):
    from oracle import detr_oracle as O
    gen = torch.Generator().manual_seed(9)
    logits = torch.randn(100, 92, generator=gen)
    boxes = torch.cat([torch.rand(100, 2, generator=gen) * 0.9 + 0.05, torch.rand(100, 2, generator=gen) * 0.4 + 0.05], -1)
    tb, tc = O.synthetic_targets(1, n=7, seed=9)
    ti, pi, tsel, psel, tbox, tcls = D.hungarian_matching(tb[0], tc[0], boxes, logits)
    oti, opi, _, _, _, _ = O.hungarian_matching(tb[0], tc[0], boxes, logits)
    assert torch.equal(ti.cpu(), oti) and torch.equal(pi.cpu(), opi) and int(psel.sum()) == 7 and tbox.shape == (7, 4)
    bad = logits.clone()
    bad[3, 5] = float("nan")
    with pytest.raises(ValueError, match="invalid numeric entries"):
        D.hungarian_matching(tb[0], tc[0], boxes, bad)
    # get_losses: the same condition poisons every returned scalar (stream-ordered, no exception possible)
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    total, losses = D.get_losses({"pred_logits": bad[None], "pred_boxes": boxes[None], "aux": []}, tb, tc, cfg)
    assert float(total) != float(total) and all(float(v) != float(v) for v in losses.values())
    total, losses = D.get_losses({"pred_logits": logits[None], "pred_boxes": boxes[None], "aux": []}, tb, tc, cfg)
    assert float(total) == float(total)


def test_checkpoint_roundtrip_and_resume_on_device(D, tmp_path):
    """SURVEY 8f N3 on the CUDA path: parameters + Adam moments + step counters survive save / load bit for bit and a resumed
    run continues identically (dropout 0: the step is deterministic up to the fp32 atomics of the weight gradients)"""
    from oracle import detr_oracle as O
    from detr_tensorflow_b200.networks import weights as Wt
    P = O.init_params(seed=4, num_encoder_layers=NE, num_decoder_layers=ND)
    img = torch.randn(1, 96, 128, 3, generator=torch.Generator().manual_seed(4))
    tb, tc = O.synthetic_targets(1, n=3, seed=4)
    cfg = _cfg(D, 1, None)
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=NE, num_decoder_layers=ND)
    opt = D.setup_optimizers(model, cfg)
    D.training.fit(model, [(img, tb, tc)] * 2, opt, cfg, 0, None)
    path = str(tmp_path / "ck.npz")
    Wt.save_checkpoint(model, path, cfg)
    cfg2 = _cfg(D, 1, None)
    model2 = D.get_detr_model(cfg2, include_top=True, dropout=0.0, num_encoder_layers=NE, num_decoder_layers=ND, weights=path, seed=123)
    Wt.load_checkpoint(model2, path, cfg2)
    e1, e2 = model.engine, model2.engine
    torch.cuda.synchronize()
    assert cfg2.global_step == 2 and torch.equal(e1.params, e2.params) and torch.equal(e1.adam_m, e2.adam_m)
    assert torch.equal(e1.adam_v, e2.adam_v) and torch.equal(e1.steps, e2.steps)
    out1, out2 = model(img, training=False), model2(img, training=False)
    assert torch.equal(out1["pred_logits"], out2["pred_logits"])
    opt2 = D.setup_optimizers(model2, cfg2)
    D.training.fit(model, [(img, tb, tc)], opt, cfg, 0, None)
    D.training.fit(model2, [(img, tb, tc)], opt2, cfg2, 0, None)
    torch.cuda.synchronize()
    assert rel(e2.params, e1.params) < 1e-6                       # identical up to the summation order of the gradient atomics


def test_fit_fused_step_equals_the_two_reference_calls(D):
    """training.fit without gradient accumulation replays ONE graph per step (run_train_and_apply_step: forward, losses, backward,
    Adam of every enabled group, weight refresh); with config.fused_optimizer_step = False it makes the reference's two calls
    (run_train_step + aggregate_grad_and_apply per group).  Same arithmetic: losses equal to rounding (weight gradients are summed
    with fp32 atomics; after an optimizer step trajectories may differ by a flipped near-tie assignment, see
    test_pipeline_gpu.py::test_uint8_frames_through_model_and_fit), the frozen group stays frozen, step counters advance alike."""
    from oracle import detr_oracle as O
    P = O.init_params(seed=5, num_encoder_layers=1, num_decoder_layers=2)
    img = torch.randn(2, 96, 128, 3, generator=torch.Generator().manual_seed(5))
    tb, tc = O.synthetic_targets(2, n=4, seed=5)
    out = {}
    for fused in (True, False):
        cfg = D.TrainingConfig()
        cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 2, None
        cfg.train_backbone, cfg.train_transformers = False, True            # the backbone group is computed but not applied
        cfg.fused_optimizer_step = fused
        model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=1, num_decoder_layers=2)
        opt = D.setup_optimizers(model, cfg)
        seen, logs = [], []
        D.training.fit(model, [(img, tb, tc)] * 4, opt, cfg, 0, None, on_step=lambda s, t, l: (seen.append(float(t)), logs.append(l)))
        eng = model.engine
        out[fused] = (seen, {k: v.clone() for k, v in eng.export_params().items()}, eng.steps[:3].cpu().tolist(), logs[-1])
    a, b = out[True], out[False]
    print("fused", a[0], "split", b[0])
    assert len(a[0]) == 4 and abs(a[0][0] - b[0][0]) <= 5e-6 * abs(b[0][0])
    assert all(abs(x - y) <= 5e-3 * abs(y) for x, y in zip(a[0], b[0])), (a[0], b[0])
    assert a[2] == b[2] and a[2][0] == 0 and a[2][1] == 4                    # Adam iterations: backbone 0, transformers 4
    for n in P:
        if n.startswith("backbone/"):
            assert torch.equal(a[1][n], b[1][n])                             # frozen group untouched in both paths
    moved = max(float((a[1][n].float() - torch.as_tensor(P[n]).float().to(a[1][n].device)).abs().max()) for n in a[1] if not n.startswith("backbone/"))
    assert moved > 0
    assert set(a[3].keys()) == set(b[3].keys())                              # same log keys (losses + learning rates)


def test_handles_own_the_device_and_a_per_thread_kernel_policy(D):
    """include/detrb.h "Handles": detrb_create checks the device, a handle carries its own switches, detrb_bind installs them for
    the calling thread only -- a thread that bound a handle with the tcgen05 GEMM switched off runs the mma.sync kernel and gets
    the same product, while the main thread's switches (and a second handle) stay untouched."""
    import threading
    from detr_tensorflow_b200 import _lib, ops
    with pytest.raises(_lib.DetrbError):
        ops.Handle(torch.cuda.device_count() + 3)
    before = {k: None for k in ("tc", "tc_stream", "tc_halo")}
    for k in before:                                    # the main thread's current values (read through the setters)
        f = getattr(ops, "set_" + k)
        before[k] = f(1)
        f(before[k])
    h1, h2 = ops.Handle(0), ops.Handle(0)
    assert h1.device_index == 0 and h1.get("tc") == before["tc"] and h1.get("tc_pair") == -1
    h1.set("tc", 0)
    h1.set("tc_stream", 0)
    assert h2.get("tc") == before["tc"] and h2.get("tc_stream") == before["tc_stream"]          # handles are independent
    with pytest.raises(KeyError):
        h1.set("no_such_option", 1)
    g = torch.Generator(device="cuda").manual_seed(3)
    M, N, K = 384, 128, 256
    A = torch.randn(M, K, device="cuda", generator=g).bfloat16()
    W = torch.randn(N, K, device="cuda", generator=g).bfloat16()
    ref = A.float() @ W.float().t()
    out, seen = {}, {}

    def worker():
        h1.bind()
        seen["tc"] = ops.set_tc(0)                      # previous value of THIS thread = the handle's
        seen["stream"] = ops.set_tc_stream(0)
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
            ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), C=C, ldc=N)
        s.synchronize()
        out["worker"] = C

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert seen == {"tc": 0, "stream": 0}
    assert ops.set_tc(before["tc"]) == before["tc"] and ops.set_tc_stream(before["tc_stream"]) == before["tc_stream"]   # main thread untouched
    C = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    ops.igemm(A, W, M, N, K, K, K, ops.plain_geom(M, K), C=C, ldc=N)
    torch.cuda.synchronize()
    assert rel(C, ref) < 4e-3 and rel(out["worker"], ref) < 4e-3
    h2.bind()                                           # binding a default handle leaves this thread on the defaults
    assert ops.set_tc(before["tc"]) == before["tc"]
    h1.close()
    h2.close()
    eng = D.get_detr_model(_cfg(D, 1, None), include_top=True, num_encoder_layers=1, num_decoder_layers=1).engine
    assert eng.handle is not None and eng.handle.device_index == torch.cuda.current_device()
