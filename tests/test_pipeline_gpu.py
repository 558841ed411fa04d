"""GPU parity of the rows either side of the train step, through the C ABI (csrc/pipeline.cu):
inference post-process and uint8 input normalisation, against vectors produced by the reference's own code
(tests/golden/make_golden_pipeline.py) and against the oracle at full size.  Integer outputs (labels, kept query ids,
counts) must be exact; normalised pixels are bit-exact by construction (table built with the reference's float64
arithmetic); scores within 1e-6 relative (fp32 summation order of the softmax denominator)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def D():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    import detr_tensorflow_b200 as D
    return D


@pytest.fixture(scope="module")
def g():
    return np.load(os.path.join(HERE, "golden", "pipeline_golden.npz"))


def test_postprocess_vs_reference_golden(D, g):
    logits, boxes = torch.from_numpy(g["i_logits"]).cuda(), torch.from_numpy(g["i_boxes"]).cuda()
    for bg in (91, 0):
        for fmt in ("xy_center", "xyxy", "yxyx"):
            for b in range(3):
                pb, pl, ps = D.get_model_inference({"pred_logits": logits[b:b + 1], "pred_boxes": boxes[b:b + 1]}, bg, fmt)
                k = f"i_{bg}_{fmt}_{b}"
                assert pl.dtype == torch.int64 and np.array_equal(pl.cpu().numpy(), g[k + "_labels"])      # exact
                np.testing.assert_allclose(ps.cpu().numpy(), g[k + "_scores"], rtol=1e-6)
                assert np.array_equal(pb.cpu().numpy(), g[k + "_bbox"])                                     # same fp32 ops: exact
            # the batched launch returns the same rows for every image
            ob, ol, os_, oq, cnt = D.inference.batched_model_inference({"pred_logits": logits, "pred_boxes": boxes}, bg, fmt)
            for b in range(3):
                k = f"i_{bg}_{fmt}_{b}"
                n = int(cnt[b])
                assert n == len(g[k + "_labels"]) and np.array_equal(ol[b, :n].cpu().numpy(), g[k + "_labels"])
                assert np.array_equal(ob[b, :n].cpu().numpy(), g[k + "_bbox"])
                q = oq[b, :n].cpu().numpy()
                assert np.all(np.diff(q) > 0)                                                               # ascending query order


def test_postprocess_full_size_vs_oracle_and_edges(D):
    """B=256 x Q=100 x C=92 (config-5 sized) against the oracle; Q not a multiple of 32; all-background and none-background"""
    from oracle import detr_oracle as O
    gen = torch.Generator().manual_seed(11)
    B, Q, C = 256, 100, 92
    logits = torch.randn(B, Q, C, generator=gen) * 3
    boxes = torch.cat([torch.rand(B, Q, 2, generator=gen), torch.rand(B, Q, 2, generator=gen) * 0.9 + 0.01], -1)
    ob, ol, os_, oq, cnt = D.inference.batched_model_inference({"pred_logits": logits.cuda(), "pred_boxes": boxes.cuda()}, 91, "yxyx")
    ob, ol, os_, cnt = ob.cpu(), ol.cpu(), os_.cpu(), cnt.cpu()
    for b in range(0, B, 17):
        rb, rl, rs = O.get_model_inference({"pred_logits": logits[b:b + 1], "pred_boxes": boxes[b:b + 1]}, 91, "yxyx")
        n = int(cnt[b])
        assert n == len(rl) and torch.equal(ol[b, :n], rl) and torch.equal(ob[b, :n], rb)
        assert float((os_[b, :n] - rs).abs().max()) <= 1e-6 * float(rs.max())
    for Q2, C2 in ((7, 4), (100, 4), (300, 1000)):
        lg = torch.randn(2, Q2, C2, generator=gen)
        bx = torch.rand(2, Q2, 4, generator=gen)
        lg[0, :, 1] += 50.0                                  # image 0: every query predicts class 1
        for bg, expect0 in ((1, 0), (0, Q2)):
            _, ol2, _, oq2, c2 = D.inference.batched_model_inference({"pred_logits": lg.cuda(), "pred_boxes": bx.cuda()}, bg, "xyxy")
            assert int(c2[0]) == expect0
            rb, rl, rs = O.get_model_inference({"pred_logits": lg[1:2], "pred_boxes": bx[1:2]}, bg, "xyxy")
            assert int(c2[1]) == len(rl) and torch.equal(ol2[1, :len(rl)].cpu(), rl)


def test_normalize_u8_vs_reference_golden_and_s2d_fusion(D, g):
    from detr_tensorflow_b200 import ops
    cfg = D.TrainingConfig()
    for method in ("torch_resnet", "tf_resnet"):
        cfg.normalized_method = method
        out = D.data.normalized_images(g["n_img"], cfg)
        assert out.is_cuda and np.array_equal(out.cpu().numpy(), g[f"n_{method}"])                          # bit-exact
        # odd pixel counts / unaligned views take the tail kernel
        one = D.data.normalized_images(torch.from_numpy(g["n_img"][1, 3:4, 1:6]).contiguous(), cfg)
        assert np.array_equal(one.cpu().numpy(), g[f"n_{method}"][1, 3:4, 1:6])
        # fused uint8 -> normalised bf16 space-to-depth == (normalise to fp32, then the fp32 layout kernel)
        u8 = torch.randint(0, 256, (2, 37, 53, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(3)).cuda()
        lut, swap = D.data.processing.device_lut(method, u8.device)
        a = torch.zeros(2, 19, 27, 16, dtype=torch.bfloat16, device="cuda")
        b = torch.ones_like(a)
        ops.image_u8_to_s2d16(u8, lut, swap, a, 2, 37, 53)
        ops.image_to_s2d16(D.data.normalized_images(u8, cfg), b, 2, 37, 53)
        assert torch.equal(a, b)


def test_uint8_frames_through_model_and_fit(D):
    """uint8 batches through model() and training.fit (graph-replayed step) == the normalised float batches"""
    from oracle import detr_oracle as O
    P = O.init_params(seed=4, num_encoder_layers=1, num_decoder_layers=2)
    cfg = D.TrainingConfig()
    cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 2, None
    cfg.train_backbone, cfg.train_transformers = True, True
    u8 = torch.randint(0, 256, (2, 96, 128, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
    f32 = torch.from_numpy(O.normalized_images(u8.numpy(), "torch_resnet"))
    tb, tc = O.synthetic_targets(2, n=4, seed=4)
    losses = {}
    for name, imgs in (("u8", u8), ("f32", f32)):
        model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=1, num_decoder_layers=2)
        out = model(imgs, training=False)["pred_logits"].clone()
        opt = D.setup_optimizers(model, cfg)
        seen = []
        D.training.fit(model, [(imgs, tb, tc)] * 3, opt, cfg, 0, None, on_step=lambda s, t, l: seen.append(float(t)))
        losses[name] = (out, seen)
    assert torch.equal(losses["u8"][0], losses["f32"][0])
    a, b = losses["u8"][1], losses["f32"][1]
    # same bf16 stem input -> same forward, bit for bit (logits above); the loss sums and the weight gradients are reduced with
    # fp32 atomics whose order varies run to run -> the first loss is equal to rounding.  After an optimizer step the trajectories
    # may separate DISCRETELY: the first Adam step is sign-like (lr * g / (|g| + eps)), so gradients that differ in their last bits
    # move near-zero weights in opposite directions, and with random-init weights the 100 queries are near-ties for the matcher
    # (costs within 1e-6): one flipped assignment changes the later losses by ~1e-3.  Measured over 24 runs of the SAME input
    # (tests/repro_fit_race.py, also on the build before this test existed): two outcomes, 16.7182 / 14.7337 and 16.7216 / 14.7544.
    print("losses u8", a, "f32", b)
    assert len(a) == 3 and abs(a[0] - b[0]) <= 5e-6 * abs(b[0]), (a, b)      # ~400 fp32 terms summed in varying order: a few 1e-7 .. 1e-6
    assert all(abs(x - y) <= 5e-3 * abs(y) for x, y in zip(a, b)), (a, b)
    assert losses["u8"][1][2] != losses["u8"][1][0]           # the optimizer moved


def test_map_matching_and_evaluation_vs_reference_golden(D):
    """the mAP path on device (map_match_kernel behind loss/compute_map.py) against the reference's own
    compute_map.py (tests/golden/map_golden.npz): per-image reference-style cal_map calls, and the batched wire-format matcher"""
    import os
    from detr_tensorflow_b200.loss import compute_map as CM
    from oracle import map_oracle as MO
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "map_golden.npz"))
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
    np.testing.assert_allclose(list(maps["box"].values()), g["box_map_values"], atol=1e-9)
    np.testing.assert_allclose(list(maps["mask"].values()), g["mask_map_values"], atol=1e-9)


def test_map_matching_batched_vs_oracle(D):
    """64 random images (up to 100 detections, up to 30 targets, 20 classes, score ties and duplicate boxes) through ONE matching
    launch in the padded wire format: true-positive flags, ranks and the summary must equal the oracle's (bit-exact integers)"""
    from detr_tensorflow_b200.loss import compute_map as CM
    from oracle import map_oracle as MO
    rs = np.random.RandomState(5)
    B, Q, ncls = 64, 100, 20
    thr = CM.IOU_THRESHOLDS
    boxes, labels, scores = np.zeros((B, Q, 4), np.float32), np.zeros((B, Q), np.int64), np.zeros((B, Q), np.float32)
    count = np.zeros(B, np.int32)
    tb, tc = np.zeros((B, 100, 4), np.float32), np.zeros((B, 100, 1), np.int64)
    oap = [[MO.APData() for _ in range(ncls)] for _ in thr]
    for b in range(B):
        n, k = int(rs.randint(0, 31)), int(rs.randint(0, Q + 1))
        t = np.concatenate([rs.uniform(0.1, 0.9, (n, 2)), rs.uniform(0.05, 0.6, (n, 2))], -1).astype(np.float32)
        tcls = rs.randint(0, ncls, n)
        if n:
            src = rs.randint(0, n, k)
            p = (t[src] + rs.normal(0, 0.02, (k, 4)) * (rs.rand(k, 1) < 0.7)).astype(np.float32)
            pcls = np.where(rs.rand(k) < 0.8, tcls[src], rs.randint(0, ncls, k))
        else:
            p = np.concatenate([rs.uniform(0.1, 0.9, (k, 2)), rs.uniform(0.05, 0.6, (k, 2))], -1).astype(np.float32)
            pcls = rs.randint(0, ncls, k)
        sc = np.round(rs.uniform(0.05, 1.0, k), 2).astype(np.float32)              # two decimals: many exact score ties
        pyx = MO.yxyx_from_xcycwh(p)
        boxes[b, :k], labels[b, :k], scores[b, :k], count[b] = pyx, pcls, sc, k
        tb[b, 0, 0] = n
        tb[b, 1:1 + n], tc[b, 1:1 + n, 0] = t, tcls
        MO.cal_map_image(pyx, pcls, sc, MO.yxyx_from_xcycwh(t), tcls, oap, thr)
    dev = "cuda"
    rank, tp, gtc = CM._match(torch.from_numpy(boxes).to(dev), torch.from_numpy(labels).to(dev), torch.from_numpy(scores).to(dev),
                              torch.from_numpy(count).to(dev), torch.from_numpy(tb).to(dev), torch.from_numpy(tc).to(dev), None, True, thr, ncls)
    ev = CM.MapEvaluator([f"c{i}" for i in range(ncls)], with_mask_rows=False)
    ev.batches.append((torch.from_numpy(labels), torch.from_numpy(scores), rank, tp, torch.from_numpy(count), torch.from_numpy(tb[:, 0, 0]),
                       torch.from_numpy(tc[:, :, 0])))
    ap = ev.ap_data()
    for a in range(len(thr)):
        for c in range(ncls):
            o, m = oap[a][c], ap["box"][a][c]
            assert o.num_gt_positives == m.num_gt_positives and len(o.data_points) == len(m.data_points)
            assert sorted(o.data_points, key=lambda x: -x[0]) == sorted(((float(s), bool(f)) for s, f in m.data_points), key=lambda x: -x[0]), (a, c)
            assert abs(o.get_ap() - m.get_ap()) < 1e-12
    expect = MO.calc_map(oap, thr, ncls)
    got = ev.summary()["box"]
    assert list(got.values()) == list(expect.values())
    expect_gt = np.zeros(ncls, np.int64)
    for b in range(B):
        expect_gt += np.bincount(tc[b, 1:1 + int(tb[b, 0, 0]), 0], minlength=ncls)
    assert gtc.cpu().tolist() == expect_gt.tolist()


def test_eval_model_runs_end_to_end(D, capsys):
    """eval.py:30-61 through the engine: forward -> device post-process -> device matching -> summary dict with the reference's keys"""
    from oracle import detr_oracle as O
    from detr_tensorflow_b200.loss import compute_map as CM
    P = O.init_params(seed=6, num_encoder_layers=1, num_decoder_layers=1)
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=1, num_decoder_layers=1)
    img = torch.randn(2, 96, 128, 3, generator=torch.Generator().manual_seed(6))
    tb, tc = O.synthetic_targets(2, n=4, seed=6)
    maps = CM.eval_model(model, cfg, [f"c{i}" for i in range(92)], [(img, tb, tc)] * 2, print_result=True)
    assert list(maps["box"].keys()) == ["all", 50, 55, 60, 65, 70, 75, 80, 85, 90, 95] and "mask" in maps
    assert "box" in capsys.readouterr().out


def test_resize_affine_kernel_bit_exact_vs_oracle(D):
    """resize_affine_u8_kernel (transformation.py:54-114 as one affine map per image) against the numpy oracle,
    bit for bit: ragged batch, flips, crops, zero-filled down-scaling, up- and down-sampling; then the benchmark geometry
    (8 frames of ~480x640 -> 800x1333) and detr_transform_batch end to end into the model's uint8 input."""
    from oracle.resize_oracle import resize_affine_u8
    g = np.random.default_rng(11)
    frames = [g.integers(0, 256, (int(g.integers(17, 120)), int(g.integers(17, 120)), 3), dtype=np.uint8) for _ in range(7)]
    H, W = 61, 83
    fwds, zbs = [], []
    for i, f in enumerate(frames):
        h, w = f.shape[:2]
        fx, fy = W / w * g.uniform(0.5, 1.5), H / h * g.uniform(0.5, 1.5)
        if i % 2:
            fwds.append((-fx, W - g.uniform(-5, 5), fy, g.uniform(-5, 5)))
        else:
            fwds.append((fx, g.uniform(-5, 5), fy, g.uniform(-5, 5)))
        zbs.append(i % 3 == 0)
    out = D.data.resample_batch(frames, fwds, zbs, (H, W)).cpu().numpy()
    inv = np.array([D.data.transformation.inverse_map(f) for f in fwds], np.float32)
    ref = resize_affine_u8(frames, inv, zbs, H, W)
    assert np.array_equal(out, ref)
    # identity and flip are exact copies
    same = D.data.resample_batch(frames[:1], [(1.0, 0.0, 1.0, 0.0)], [False], frames[0].shape[:2]).cpu().numpy()[0]
    assert np.array_equal(same, frames[0])
    w0 = frames[0].shape[1]
    flip = D.data.resample_batch(frames[:1], [(-1.0, float(w0), 1.0, 0.0)], [False], frames[0].shape[:2]).cpu().numpy()[0]
    assert np.array_equal(flip, frames[0][:, ::-1])
    # benchmark geometry, through the public function; one image checked against the oracle
    cfg = D.TrainingConfig()
    cfg.image_size = (800, 1333)
    big = [g.integers(0, 256, (480 + 8 * i, 640 - 8 * i, 3), dtype=np.uint8) for i in range(8)]
    boxes = [np.array([[0.5, 0.5, 0.3, 0.3], [0.2, 0.7, 0.1, 0.2]]) for _ in big]
    cls = [np.array([1, 2]) for _ in big]
    batch, nb, nc = D.data.detr_transform_batch(big, boxes, cls, cfg, False)
    assert batch.dtype == torch.uint8 and tuple(batch.shape) == (8, 800, 1333, 3) and batch.is_cuda
    ref3 = resize_affine_u8([big[3]], np.array([[big[3].shape[1] / 1333, 0, big[3].shape[0] / 800, 0]], np.float32), [0], 800, 1333)[0]
    assert np.array_equal(batch[3].cpu().numpy(), ref3)
    np.testing.assert_allclose(nb[3], boxes[3], atol=1e-12)
    aug, _, _ = D.data.detr_transform_batch(big, boxes, cls, cfg, True, rng=np.random.default_rng(1))
    assert tuple(aug.shape) == (8, 800, 1333, 3)
