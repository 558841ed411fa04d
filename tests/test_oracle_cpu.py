"""CPU tests pinning the oracle: (1) against golden vectors produced by executing the reference's own
loss/matcher code (tests/golden/make_golden.py), (2) the C LSAP restatement against the installed scipy
(the reference's real third-party arithmetic) incl. the tie known-answer table of SURVEY Appendix C,
(3) model pieces against independent implementations (torchvision resnet50, F.multi_head_attention_forward)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import detr_oracle as O
from oracle.lsap_oracle import lsap


def test_cost_matrix_and_indices_vs_reference_golden(golden):
    g = golden
    B = g["m_logits"].shape[0]
    for b in range(B):
        tb, tc = torch.from_numpy(g["m_t_bbox"][b]), torch.from_numpy(g["m_t_class"][b])
        pb, pc = torch.from_numpy(g["m_boxes"][b]), torch.from_numpy(g["m_logits"][b])
        C, _, _ = O.cost_matrix(tb, tc, pb, pc)
        np.testing.assert_allclose(C.numpy(), g[f"m_cost_{b}"], rtol=1e-5, atol=2e-6)
        ti, pi, ts, ps, tbb, tcc = O.hungarian_matching(tb, tc, pb, pc)
        assert np.array_equal(ti.numpy(), g[f"m_t_indices_{b}"])
        assert np.array_equal(pi.numpy(), g[f"m_p_indices_{b}"])
        assert np.array_equal(ps.numpy(), g[f"m_p_selector_{b}"])
        assert np.array_equal(ts.numpy(), g[f"m_t_selector_{b}"])
        np.testing.assert_array_equal(tbb.numpy(), g[f"m_tb_{b}"])
        np.testing.assert_array_equal(tcc.numpy(), g[f"m_tc_{b}"])
        # the C restatement on the reference's own cost matrix
        r, c = lsap(g[f"m_cost_{b}"])
        assert np.array_equal(r, g[f"m_p_indices_{b}"]) and np.array_equal(c, g[f"m_t_indices_{b}"])


def test_set_criterion_vs_reference_golden(golden):
    g = golden
    logits, boxes = torch.from_numpy(g["l_logits"]), torch.from_numpy(g["l_boxes"])
    out = {"pred_logits": logits[5], "pred_boxes": boxes[5],
           "aux": [{"pred_logits": logits[i], "pred_boxes": boxes[i]} for i in range(5)]}
    total, losses = O.get_losses(out, torch.from_numpy(g["l_t_bbox"]), torch.from_numpy(g["l_t_class"]), 91)
    assert abs(float(total) - float(g["l_total"])) < 2e-4 * abs(float(g["l_total"]))
    for k, v in zip(g["l_keys"], g["l_values"]):
        assert abs(float(losses[str(k)]) - float(v)) < 1e-5 + 1e-4 * abs(float(v)), k
    assert len(losses) == 36


KAT = [  # SURVEY.md Appendix C (scipy 1.18.1)
    ([[1], [1], [.5], [.5]], [2], [0]),
    ([[1, 2], [1, 2], [.5, 3], [.5, 3]], [0, 2], [1, 0]),
    (np.ones((6, 2)), [0, 1], [0, 1]),
    (np.ones((3, 3)), [0, 1, 2], [0, 1, 2]),
    ([[0, 0], [1, 1], [2, 2], [3, 3], [4, 4]], [0, 1], [0, 1]),
    ([[3, 3], [1, 1], [1, 1], [2, 2], [5, 5]], [1, 2], [0, 1]),
    (np.zeros((2, 5)), [0, 1], [0, 1]),
    ([[1, np.inf], [2, 3], [np.inf, 1]], None, None),
]


def test_lsap_known_answers():
    from scipy.optimize import linear_sum_assignment
    for cost, rows, cols in KAT:
        c = np.asarray(cost, np.float32)
        r0, c0 = linear_sum_assignment(c)
        r1, c1 = lsap(c)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1)
        if rows is not None:
            assert list(r1) == rows and list(c1) == cols
    r, c = lsap(np.zeros((100, 0), np.float32))
    assert len(r) == 0 and len(c) == 0
    with pytest.raises(ValueError):
        lsap(np.array([[1, np.nan], [0, 1]], np.float32))


def test_lsap_vs_scipy_random_and_ties():
    from scipy.optimize import linear_sum_assignment
    rs = np.random.RandomState(7)
    for it in range(4000):
        n = rs.randint(0, 100) if it % 2 else rs.randint(0, 30)
        if it % 3 == 0:
            c = rs.rand(100, n).astype(np.float32)
        elif it % 3 == 1:
            c = rs.randint(0, 4, (100, n)).astype(np.float32)
        else:
            c = (rs.randint(0, 3, (100, n)) * 0.25).astype(np.float32) - rs.randint(0, 2, (100, 1))
        r0, c0 = linear_sum_assignment(c)
        r1, c1 = lsap(c)
        assert np.array_equal(r0, r1) and np.array_equal(c0, c1)


def test_backbone_vs_torchvision():
    tv = pytest.importorskip("torchvision")
    m = tv.models.resnet50(weights=None).eval()
    P = O.init_params(seed=3)
    sd = m.state_dict()

    def put(conv, name):
        sd[conv + ".weight"] = P[name].permute(3, 2, 0, 1).contiguous()

    def putbn(bn, name):
        sd[bn + ".weight"], sd[bn + ".bias"] = P[name + "/weight"], P[name + "/bias"]
        sd[bn + ".running_mean"], sd[bn + ".running_var"] = P[name + "/running_mean"], P[name + "/running_var"]
    put("conv1", "backbone/conv1/kernel")
    putbn("bn1", "backbone/bn1")
    for li, (nb, _, _, _) in enumerate(O.RESNET_STAGES["resnet50"]):
        for b in range(nb):
            t, p = f"layer{li + 1}.{b}", f"backbone/layer{li + 1}/{b}"
            for k in (1, 2, 3):
                put(f"{t}.conv{k}", f"{p}/conv{k}/kernel")
                putbn(f"{t}.bn{k}", f"{p}/bn{k}")
            if b == 0:
                put(f"{t}.downsample.0", f"{p}/downsample_0/kernel")
                putbn(f"{t}.downsample.1", f"{p}/downsample_1")
    m.load_state_dict(sd)
    x = torch.randn(1, 67, 93, 3)
    with torch.no_grad():
        y = O.backbone_forward(P, x)
        xt = x.permute(0, 3, 1, 2)
        z = m.layer4(m.layer3(m.layer2(m.layer1(m.maxpool(m.relu(m.bn1(m.conv1(xt))))))))
    assert y.shape == (1, 3, 3, 2048)
    torch.testing.assert_close(y.permute(0, 3, 1, 2), z, rtol=2e-4, atol=2e-4)


def test_backbone_shapes_800x1333_formula():
    # spatial sizes of SURVEY 8: 800x1333 -> 400x667 -> 200x334 -> 100x167 -> 50x84 -> 25x42
    def o(n, k, s, p):
        return (n + 2 * p - k) // s + 1
    h, w = o(800, 7, 2, 3), o(1333, 7, 2, 3)
    assert (h, w) == (400, 667)
    h, w = o(h, 3, 2, 1), o(w, 3, 2, 1)
    assert (h, w) == (200, 334)
    for exp in ((100, 167), (50, 84), (25, 42)):
        h, w = o(h, 3, 2, 1), o(w, 3, 2, 1)
        assert (h, w) == exp


def test_mha_vs_torch_functional():
    P = O.init_params(seed=5, num_encoder_layers=1, num_decoder_layers=1)
    p = "transformer/encoder/layer_0/self_attn"
    q, k, v = torch.randn(3, 11, 256), torch.randn(3, 37, 256), torch.randn(3, 37, 256)
    a = O.mha(P, p, q, k, v)
    b, _ = F.multi_head_attention_forward(
        q.transpose(0, 1), k.transpose(0, 1), v.transpose(0, 1), 256, 8,
        P[p + "/in_proj_kernel"], P[p + "/in_proj_bias"], None, None, False, 0.0,
        P[p + "/out_proj_kernel"], P[p + "/out_proj_bias"], training=False, need_weights=False)
    torch.testing.assert_close(a, b.transpose(0, 1), rtol=1e-4, atol=1e-5)


def test_forward_shapes_and_param_count():
    P = O.init_params(seed=0)
    assert sum(p.numel() for n, p in P.items() if O.param_group(n)) == 41499168   # SURVEY 8a totals
    with torch.no_grad():
        out = O.detr_forward(P, torch.randn(1, 64, 96, 3))
    assert out["pred_logits"].shape == (1, 100, 92) and out["pred_boxes"].shape == (1, 100, 4)
    assert len(out["aux"]) == 5


@pytest.mark.parametrize("case", ["a", "b", "ft"])
def test_model_forward_vs_reference_code_golden(case):
    """The MODEL part of the oracle against tests/golden/model_golden.npz: activations produced by the reference's own
    detr.py / resnet_backbone.py / transformer.py / custom_layers.py / position_embeddings.py, executed unmodified on a
    torch-backed TensorFlow shim (tests/golden/make_golden_model.py) with these same seeded weights injected by Keras
    variable name.  Pins layer wiring, padding / stride placement, the packed in-projection split, the query scaling, the
    [pos_y, pos_x] sin/cos interleave, the [S,B,256] <-> NHWC transposes and the head stack.  fp32 on both sides:
    tolerance = accumulation-order noise.  Case "ft" goes through the fine-tuning heads (detr.py:94-114: Keras Dense
    `cls_layer` + Sequential `pos_layer`, variables pos_layer/dense{,_1,_2}/...)."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz"))
    seed, B, H, W, ne, nd, nb_class = (int(v) for v in g[f"{case}_meta"])
    nb_class = nb_class or None                                        # "ft": include_top=False, nb_class=3 -> add_heads_nlayers
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd, nb_class=nb_class)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    # the reference created exactly these variables (names = Keras layer-name paths) and marks these trainable
    assert sorted(n for n in P if not n.split("/")[-1] in ("weight", "bias", "running_mean", "running_var")
                  or not ("bn" in n.split("/")[-2] or n.split("/")[-2] == "downsample_1")) == sorted(g[f"{case}_trainable"].tolist())
    with torch.no_grad():
        feat = O.backbone_forward(P, img)
        out = O.detr_forward(P, img, num_encoder_layers=ne, num_decoder_layers=nd, return_hs=True)
        pos = O.position_embedding_sine(feat.shape[1], feat.shape[2])

    def close(name, got, ref, tol):
        got, ref = torch.as_tensor(got).float(), torch.from_numpy(np.asarray(ref)).float()
        assert got.shape == ref.shape, (name, got.shape, ref.shape)
        err = float((got - ref).abs().max()) / (float(ref.abs().max()) + 1e-12)
        assert err < tol, (name, err)
    close("backbone", feat, g[f"{case}_feat"], 2e-5)
    close("position embedding", pos.reshape(1, feat.shape[1], feat.shape[2], 256).expand(B, -1, -1, -1), g[f"{case}_pos"], 1e-5)
    close("hs", out["hs"], g[f"{case}_hs"], 5e-5)
    close("pred_logits", out["pred_logits"], g[f"{case}_pred_logits"], 5e-5)
    close("pred_boxes", out["pred_boxes"], g[f"{case}_pred_boxes"], 5e-5)
    assert len(out["aux"]) == nd - 1
    for i, a in enumerate(out["aux"]):
        close(f"aux{i} logits", a["pred_logits"], g[f"{case}_aux{i}_logits"], 5e-5)
        close(f"aux{i} boxes", a["pred_boxes"], g[f"{case}_aux{i}_boxes"], 5e-5)


def test_resnet101_backbone_vs_reference_code_golden():
    """resnet_backbone.py:52-66 (ResNet101Backbone, the C4 configuration's backbone) executed on the shim vs the oracle"""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz"))
    seed, B, H, W = (int(v) for v in g["r101_meta"])
    P = O.init_params(seed=seed, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    with torch.no_grad():
        feat = O.backbone_forward(P, img, backbone="resnet101")
    ref = torch.from_numpy(g["r101_feat"])
    assert feat.shape == ref.shape and float((feat - ref).abs().max()) < 5e-5 * float(ref.abs().max())


def test_train_step_gradients_vs_reference_code_golden():
    """training.py:9-25 restated (forward -> get_losses -> gradients of every trainable variable) against the gradient of the
    REFERENCE'S OWN loss code (loss.py / hungarian_matching.py / bbox.py, real scipy) through the REFERENCE'S OWN model code,
    taken with torch.autograd on the TensorFlow shim (tests/golden/make_golden_model.py::train_case).  Every trainable
    variable: gradient norm, a seeded random projection, and the full tensor for the 42 small ones."""
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "model_golden.npz"))
    seed, B, H, W, ne, nd, n_t = (int(v) for v in g["train_meta"])
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    tb, tc = torch.from_numpy(g["train_t_bbox"]), torch.from_numpy(g["train_t_class"])
    match = torch.from_numpy(g["train_match"])
    # the oracle's own Hungarian step finds the reference's assignment ...
    out, total, log, grads = O.train_step(P, img, tb, tc, num_encoder_layers=ne, num_decoder_layers=nd)
    _, idx = O.get_detr_losses(out["pred_logits"], out["pred_boxes"], tb, tc, 91, return_indices=True)
    _, idx_ref = O.get_detr_losses(out["pred_logits"], out["pred_boxes"], tb, tc, 91, return_indices=True, match_override=match[nd - 1])
    assert all(torch.equal(a, b) for a, b in zip(idx, idx_ref))
    # ... and under it the same losses and gradients
    assert abs(float(total) - float(g["train_total"])) < 2e-5 * abs(float(g["train_total"]))
    for k, v in zip(g["train_loss_keys"].tolist(), g["train_loss_values"].tolist()):
        assert abs(float(log[k]) - v) < 2e-5 * max(abs(v), 1e-3), (k, float(log[k]), v)
    names = g["train_names"].tolist()
    # variables the reference marks trainable == the oracle's three optimizer groups + query_embed (trainable, but called
    # outside the functional graph, detr.py:175 -> reached by no optimizer: optimizers.py:10-43)
    assert set(names) - set(grads) == {"query_embed/kernel"} and set(grads) <= set(names)
    for i, n in enumerate(names):
        if n not in grads:
            continue
        gr = grads[n].float()
        ref_norm, ref_proj = float(g["train_grad_norms"][i]), float(g["train_grad_projs"][i])
        r = torch.randn(gr.shape, generator=torch.Generator().manual_seed(1000 + i))
        assert abs(float(gr.norm()) - ref_norm) <= 1e-3 * ref_norm + 1e-9, (n, float(gr.norm()), ref_norm)
        assert abs(float((gr * r).sum()) - ref_proj) <= 1e-3 * ref_norm * gr.numel() ** 0.5 + 1e-9, (n, float((gr * r).sum()), ref_proj)
        if "train_grad/" + n in g:
            full = torch.from_numpy(g["train_grad/" + n])
            assert float((gr - full).abs().max()) <= 1e-3 * float(full.abs().max()) + 1e-9, n


def test_baseline_config_c1_forward_480x640_cpu():
    """BASELINE.json configs[0]: DETR-R50 forward on 1 synthetic 480x640 image on the CPU (plumbing, no GPU): the reference-shaped
    output dict with [1,100,92] logits, [1,100,4] boxes in [0,1] and 5 aux entries; feature map 15x20 (S = 300)."""
    P = O.init_params(seed=0)
    img = torch.randn(1, 480, 640, 3, generator=torch.Generator().manual_seed(0))
    with torch.no_grad():
        feat = O.backbone_forward(P, img)
        out = O.detr_forward(P, img)
    assert tuple(feat.shape[1:3]) == (15, 20)
    assert out["pred_logits"].shape == (1, 100, 92) and out["pred_boxes"].shape == (1, 100, 4) and len(out["aux"]) == 5
    assert all(a["pred_logits"].shape == (1, 100, 92) and a["pred_boxes"].shape == (1, 100, 4) for a in out["aux"])
    assert bool(torch.isfinite(out["pred_logits"]).all()) and float(out["pred_boxes"].min()) >= 0 and float(out["pred_boxes"].max()) <= 1


def test_pos_embedding_layout():
    pe = O.position_embedding_sine(3, 4)
    assert pe.shape == (3, 4, 256)
    # first 128 channels depend on the row only, last 128 on the column only (concat [pos_y, pos_x])
    assert torch.allclose(pe[0, 0, :128], pe[0, 3, :128]) and torch.allclose(pe[0, 1, 128:], pe[2, 1, 128:])
    y = (1.0 / (3 + 1e-6)) * 2 * np.pi
    assert abs(float(pe[0, 0, 0]) - np.sin(y)) < 1e-6 and abs(float(pe[0, 0, 1]) - np.cos(y)) < 1e-6


# ------------------------------------------------------------------------------------------------ N1 / N2 rows
@pytest.fixture(scope="module")
def pipeline_golden():
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pipeline_golden.npz"))


def test_oracle_inference_postprocess_vs_reference_golden(pipeline_golden):
    """oracle get_model_inference against the reference's own inference.py:68-95 outputs (make_golden_pipeline.py)"""
    g = pipeline_golden
    logits, boxes = torch.from_numpy(g["i_logits"]), torch.from_numpy(g["i_boxes"])
    for bg in (91, 0):
        for fmt in ("xy_center", "xyxy", "yxyx"):
            for b in range(logits.shape[0]):
                pb, pl, ps = O.get_model_inference({"pred_logits": logits[b:b + 1], "pred_boxes": boxes[b:b + 1]}, bg, fmt)
                k = f"i_{bg}_{fmt}_{b}"
                assert np.array_equal(pl.numpy(), g[k + "_labels"])                       # integer: exact
                np.testing.assert_allclose(ps.numpy(), g[k + "_scores"], rtol=1e-6, atol=1e-7)
                np.testing.assert_allclose(pb.numpy(), g[k + "_bbox"], rtol=0, atol=1e-7)


def test_oracle_normalize_and_pad_labels_vs_reference_golden(pipeline_golden):
    g = pipeline_golden
    for method in ("torch_resnet", "tf_resnet"):
        assert np.array_equal(O.normalized_images(g["n_img"], method), g[f"n_{method}"])   # same float64 arithmetic: bit-exact
    for k in range(4):
        tb, tc = O.pad_labels(torch.from_numpy(g[f"p_{k}_in_bbox"]), torch.from_numpy(g[f"p_{k}_in_class"])[:, 0])
        assert np.array_equal(tb.numpy(), g[f"p_{k}_bbox"]) and np.array_equal(tc.numpy(), g[f"p_{k}_class"])


def test_map_oracle_vs_reference_code_golden():
    """oracle/map_oracle.py against the reference's own compute_map.py (tests/golden/make_golden_map.py -> map_golden.npz):
    per (IoU threshold, class) AP, ground-truth counts, number of detections, and the rounded summary dict"""
    import os
    import numpy as np
    from oracle import map_oracle as M
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "map_golden.npz"))
    ncls, nimg = int(g["ncls"]), int(g["nimg"])
    thr = [x / 100. for x in range(50, 100, 5)]
    ap = [[M.APData() for _ in range(ncls)] for _ in thr]
    for i in range(nimg):
        M.cal_map_image(M.yxyx_from_xcycwh(g[f"p_bbox_{i}"]), g[f"p_cls_{i}"], g[f"p_score_{i}"], M.yxyx_from_xcycwh(g[f"t_bbox_{i}"]),
                        g[f"t_cls_{i}"], ap, thr)
    for a in range(len(thr)):
        for c in range(ncls):
            assert ap[a][c].num_gt_positives == g["box_ngt"][a, c] and len(ap[a][c].data_points) == g["box_npts"][a, c]
            if g["box_ap"][a, c] >= 0:
                assert abs(ap[a][c].get_ap() - g["box_ap"][a, c]) < 1e-12, (a, c)
            else:
                assert ap[a][c].is_empty()
    maps = M.calc_map(ap, thr, ncls)
    assert [str(k) for k in maps.keys()] == g["box_map_keys"].tolist()
    np.testing.assert_allclose(list(maps.values()), g["box_map_values"], atol=1e-9)


def test_oracle_resize_vs_torch_interpolate_and_box_geometry():
    """SURVEY 8f N1 (transformation.py:54-114, 163-195).  imgaug is not installed: the bilinear pixel-centre convention of the
    oracle's resampler is pinned against torch.nn.functional.interpolate(align_corners=False) (same convention as
    cv2.INTER_LINEAR, float arithmetic) to +-1 grey level, up- and down-scaling; flips, crops and zero fill by construction."""
    import torch
    import torch.nn.functional as F
    from oracle.resize_oracle import resize_affine_u8, transform_boxes
    rng = np.random.default_rng(0)
    for (h, w, H, W) in ((37, 53, 64, 80), (120, 90, 48, 64), (33, 47, 33, 47)):
        img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        out = resize_affine_u8([img], np.array([[w / W, 0.0, h / H, 0.0]], np.float32), [0], H, W)[0]
        ref = F.interpolate(torch.from_numpy(img).permute(2, 0, 1)[None].float(), size=(H, W), mode="bilinear", align_corners=False)
        ref = ref[0].permute(1, 2, 0).round().clamp(0, 255).numpy()
        assert np.abs(out.astype(np.int32) - ref.astype(np.int32)).max() <= 1
        if (h, w) == (H, W):
            assert np.array_equal(out, img)                                                 # identity map: exact
    img = rng.integers(0, 256, (20, 30, 3), dtype=np.uint8)
    flip = resize_affine_u8([img], np.array([[-1.0, 30.0, 1.0, 0.0]], np.float32), [0], 20, 30)[0]
    assert np.array_equal(flip, img[:, ::-1])                                               # Fliplr
    crop = resize_affine_u8([img], np.array([[1.0, 5.0, 1.0, 3.0]], np.float32), [0], 10, 12)[0]
    assert np.array_equal(crop, img[3:13, 5:17])                                            # CropToFixedSize at offset (5, 3)
    small = resize_affine_u8([img], np.array([[2.0, -15.0, 2.0, -10.0]], np.float32), [1], 20, 30)[0]
    assert small[0, 0].sum() == 0 and small[19, 29].sum() == 0 and small[10, 15].sum() > 0   # Affine scale 0.5: zero fill around
    # boxes: flip + scale, the 0.7 out-of-image rule, clipping
    bb = np.array([[0.25, 0.5, 0.2, 0.4], [0.95, 0.5, 0.4, 0.2], [0.995, 0.5, 0.2, 0.2]])
    nb, nc = transform_boxes(bb, np.array([3, 4, 5]), (-2.0, 200.0, 2.0, 0.0), (100, 200), (50, 100))
    np.testing.assert_allclose(nb[0], [0.75, 0.5, 0.2, 0.4], atol=1e-12)
    assert list(nc) == [3, 4, 5]
    nb, nc = transform_boxes(bb, np.array([3, 4, 5]), (2.0, 0.0, 2.0, 0.0), (100, 150), (50, 100))     # right part cut off
    assert list(nc) == [3] and nb.shape == (1, 4)                                                      # both others >= 70 % outside
    nb, nc = transform_boxes(bb[1:2], np.array([4]), (1.0, 0.0, 1.0, 0.0), (50, 100), (50, 100))       # 37.5 % outside: kept, clipped
    np.testing.assert_allclose(nb[0], [0.875, 0.5, 0.25, 0.2], atol=1e-12)
