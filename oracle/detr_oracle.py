"""CPU oracle for the DETR train-step hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This file restates, in plain PyTorch-CPU ops (fp32 or fp64), the arithmetic of the
reference's hot path (Visual-Behavior/detr-tensorflow).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline legs may import it; the
product path (``detr_tensorflow_b200``) never does.

PARITY STATUS: the reference ships no tests / golden vectors and TensorFlow is not
installable here.  What pins this oracle instead is the reference's OWN code, executed
unmodified on stand-ins for the TensorFlow library:
  * model forward (backbone, position embedding, transformer, heads): detr_tf/networks/
    {detr,resnet_backbone,transformer,custom_layers,position_embeddings}.py run through
    ``get_detr_model()`` on a torch-backed ``tensorflow`` shim with this oracle's seeded
    weights injected by Keras variable name (tests/golden/make_golden_model.py ->
    model_golden.npz); the oracle reproduces every captured activation to ~1e-5
    (tests/test_oracle_cpu.py::test_model_forward_vs_reference_code_golden), and is also
    cross-checked against torchvision resnet50 and F.multi_head_attention_forward;
  * train-step gradients: the same shim with autograd on -- the gradient of the reference's
    own get_losses through the reference's own model w.r.t. every trainable variable
    (make_golden_model.py::train_case); train_step() below matches norm / projection / small
    tensors of all of them (test_train_step_gradients_vs_reference_code_golden);
  * loss / matcher: detr_tf/loss/*.py + detr_tf/bbox.py on a numpy-backed shim
    (tests/golden/make_golden.py), with the real ``scipy.optimize.linear_sum_assignment``
    (the routine the reference calls, hungarian_matching.py:7,29);
  * inference post-process / input normalisation / pad_labels: make_golden_pipeline.py.
Still "parity unpinned" (TF library arithmetic, restated from its documentation): Keras
Adam with per-variable clipnorm, and the dropout random stream (parity is defined with
dropout off).

Every function cites the reference file:line it follows (paths relative to
/root/reference/detr_tf).
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch
import torch.nn.functional as F

# --------------------------------------------------------------------------------------
# parameter inventory
# --------------------------------------------------------------------------------------

RESNET_STAGES = {
    # networks/resnet_backbone.py:35-66  (num_bottlenecks, dim1, dim2, stride)
    "resnet50": [(3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)],
    "resnet101": [(3, 64, 256, 1), (4, 128, 512, 2), (23, 256, 1024, 2), (3, 512, 2048, 2)],
}


def param_shapes(num_classes=92, backbone="resnet50", num_encoder_layers=6,
                 num_decoder_layers=6, model_dim=256, ffn_dim=2048, num_queries=100, nb_class=None):
    """Ordered {name: (shape, kind)}; kind in {conv, bn_w, bn_b, bn_mean, bn_var, linear_w,
    linear_b, ln_g, ln_b, embed}.  Layouts are the reference's: Conv2D kernels HWIO
    (Keras), Linear kernels [out, in] (custom_layers.py:41-47), packed in_proj [3d, d]
    (transformer.py:250-268)."""
    s = OrderedDict()

    def bn(prefix, c):
        s[prefix + "/weight"] = ((c,), "bn_w")
        s[prefix + "/bias"] = ((c,), "bn_b")
        s[prefix + "/running_mean"] = ((c,), "bn_mean")
        s[prefix + "/running_var"] = ((c,), "bn_var")

    # stem  resnet_backbone.py:11-17
    s["backbone/conv1/kernel"] = ((7, 7, 3, 64), "conv")
    bn("backbone/bn1", 64)
    cin = 64
    for li, (nb, d1, d2, stride) in enumerate(RESNET_STAGES[backbone]):
        for b in range(nb):
            p = f"backbone/layer{li + 1}/{b}"
            s[p + "/conv1/kernel"] = ((1, 1, cin, d1), "conv")
            bn(p + "/bn1", d1)
            s[p + "/conv2/kernel"] = ((3, 3, d1, d1), "conv")
            bn(p + "/bn2", d1)
            s[p + "/conv3/kernel"] = ((1, 1, d1, d2), "conv")
            bn(p + "/bn3", d2)
            if b == 0:  # resnet_backbone.py:80-81 downsample=True on block 0 of EVERY stage
                s[p + "/downsample_0/kernel"] = ((1, 1, cin, d2), "conv")
                bn(p + "/downsample_1", d2)
            cin = d2
    d = model_dim
    s["input_proj/kernel"] = ((1, 1, cin, d), "conv")       # detr.py:44 (Conv2D with bias)
    s["input_proj/bias"] = ((d,), "linear_b")

    def mha(p):
        s[p + "/in_proj_kernel"] = ((3 * d, d), "linear_w")
        s[p + "/in_proj_bias"] = ((3 * d,), "linear_b")
        s[p + "/out_proj_kernel"] = ((d, d), "linear_w")
        s[p + "/out_proj_bias"] = ((d,), "linear_b")

    def lin(p, o, i):
        s[p + "/kernel"] = ((o, i), "linear_w")
        s[p + "/bias"] = ((o,), "linear_b")

    def ln(p):
        s[p + "/gamma"] = ((d,), "ln_g")
        s[p + "/beta"] = ((d,), "ln_b")

    for l in range(num_encoder_layers):
        p = f"transformer/encoder/layer_{l}"
        mha(p + "/self_attn")
        lin(p + "/linear1", ffn_dim, d)
        lin(p + "/linear2", d, ffn_dim)
        ln(p + "/norm1")
        ln(p + "/norm2")
    for l in range(num_decoder_layers):
        p = f"transformer/decoder/layer_{l}"
        mha(p + "/self_attn")
        mha(p + "/multihead_attn")
        lin(p + "/linear1", ffn_dim, d)
        lin(p + "/linear2", d, ffn_dim)
        ln(p + "/norm1")
        ln(p + "/norm2")
        ln(p + "/norm3")
    ln("transformer/decoder/norm")
    s["query_embed/kernel"] = ((num_queries, d), "embed")
    if nb_class is None:
        lin("class_embed", num_classes, d)
        lin("bbox_embed_0", d, d)
        lin("bbox_embed_1", d, d)
        lin("bbox_embed_2", 4, d)
    else:
        # add_heads_nlayers (detr.py:94-114): Keras Dense layers, kernel [in, out] (x @ kernel + bias);
        # pos_layer is a Sequential of three Dense layers (relu, relu, sigmoid)
        for p, i, o in (("cls_layer", d, nb_class), ("pos_layer/dense", d, d), ("pos_layer/dense_1", d, d),
                        ("pos_layer/dense_2", d, 4)):
            s[p + "/kernel"] = ((i, o), "dense_w")
            s[p + "/bias"] = ((o,), "linear_b")
    return s


def init_params(seed=0, dtype=torch.float32, stable=True, **kw):
    """Seeded synthetic weights (no checkpoint is reachable offline: networks/weights.py:5-11).

    stable=True : He-style conv init, BN weight~1 / bias~0, var~1: keeps activations O(1)
                  through 50 layers (SURVEY 8d, C2 note).
    stable=False: the reference's own initialisers (Glorot-uniform everywhere, incl. the
                  BN vectors, custom_layers.py:11-18) -- activations collapse; only useful
                  for shape tests.
    """
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, (shape, kind) in param_shapes(**kw).items():
        if kind == "conv":
            kh, kw_, ci, co = shape
            fan_in, fan_out = kh * kw_ * ci, kh * kw_ * co
            if stable:
                std = math.sqrt(2.0 / fan_in)
                # last conv of a residual branch a bit smaller so the residual sum stays O(1)
                if name.endswith("conv3/kernel"):
                    std *= 0.5
                if name.startswith("input_proj"):
                    std = math.sqrt(1.0 / fan_in)
                t = torch.randn(shape, generator=g, dtype=torch.float64) * std
            else:
                lim = math.sqrt(6.0 / (fan_in + fan_out))
                t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif kind == "bn_w":
            t = 1.0 + 0.1 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_b":
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_mean":
            t = 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "bn_var":
            t = 1.0 + 0.1 * torch.rand(shape, generator=g, dtype=torch.float64)
        elif kind in ("linear_w", "embed", "dense_w"):
            o, i = shape
            lim = math.sqrt(6.0 / (o + i))      # Glorot uniform, custom_layers.py:43-44 (Keras Dense default too)
            t = (torch.rand(shape, generator=g, dtype=torch.float64) * 2 - 1) * lim
        elif kind == "linear_b":
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "ln_g":
            t = 1.0 + 0.05 * torch.randn(shape, generator=g, dtype=torch.float64)
        elif kind == "ln_b":
            t = 0.02 * torch.randn(shape, generator=g, dtype=torch.float64)
        else:
            raise ValueError(kind)
        out[name] = t.to(dtype)
    return out


# --------------------------------------------------------------------------------------
# model forward
# --------------------------------------------------------------------------------------

def frozen_bn(x_nchw, P, prefix, eps=1e-5):
    """custom_layers.py:21-24"""
    scale = P[prefix + "/weight"] * torch.rsqrt(P[prefix + "/running_var"] + eps)
    shift = P[prefix + "/bias"] - P[prefix + "/running_mean"] * scale
    return x_nchw * scale.view(1, -1, 1, 1) + shift.view(1, -1, 1, 1)


def conv_hwio(x_nchw, k_hwio, stride=1, pad=0, bias=None):
    """Keras Conv2D(padding='valid') after an explicit symmetric ZeroPadding2D(pad)
    (resnet_backbone.py:11-12, 98-104).  Symmetric explicit pad + valid == torch padding=pad."""
    w = k_hwio.permute(3, 2, 0, 1)
    return F.conv2d(x_nchw, w, bias=bias, stride=stride, padding=pad)


def bottleneck(x, P, p, stride, downsample):
    """resnet_backbone.py:116-136 (stride sits on the 3x3 conv2, :104)"""
    identity = x
    out = F.relu(frozen_bn(conv_hwio(x, P[p + "/conv1/kernel"]), P, p + "/bn1"))
    out = F.relu(frozen_bn(conv_hwio(out, P[p + "/conv2/kernel"], stride=stride, pad=1), P, p + "/bn2"))
    out = frozen_bn(conv_hwio(out, P[p + "/conv3/kernel"]), P, p + "/bn3")
    if downsample:
        identity = frozen_bn(conv_hwio(x, P[p + "/downsample_0/kernel"], stride=stride), P,
                             p + "/downsample_1")
    return F.relu(out + identity)


def backbone_forward(P, images_nhwc, backbone="resnet50", return_stages=False):
    """resnet_backbone.py:20-32.  in [B,H,W,3] -> out [B,h,w,2048] (NHWC like the reference)."""
    x = images_nhwc.permute(0, 3, 1, 2)
    x = F.relu(frozen_bn(conv_hwio(x, P["backbone/conv1/kernel"], stride=2, pad=3), P, "backbone/bn1"))
    stages = [x]
    # ZeroPadding2D(1) + MaxPool(3, 2, valid): zero pad == -inf pad because x >= 0 (post-ReLU)
    x = F.max_pool2d(F.pad(x, (1, 1, 1, 1), value=0.0), 3, 2)
    stages.append(x)
    for li, (nb, d1, d2, stride) in enumerate(RESNET_STAGES[backbone]):
        for b in range(nb):
            x = bottleneck(x, P, f"backbone/layer{li + 1}/{b}", stride if b == 0 else 1, b == 0)
        stages.append(x)
    y = x.permute(0, 2, 3, 1)
    return (y, stages) if return_stages else y


def position_embedding_sine(h, w, num_pos_features=128, temperature=10000.0, eps=1e-6,
                            dtype=torch.float32):
    """position_embeddings.py:23-50 with an all-False mask (detr.py:172) -> [h, w, 256]."""
    scale = 2 * math.pi
    y_embed = torch.arange(1, h + 1, dtype=dtype).view(h, 1).expand(h, w)
    x_embed = torch.arange(1, w + 1, dtype=dtype).view(1, w).expand(h, w)
    y_embed = y_embed / (y_embed[-1:, :] + eps) * scale
    x_embed = x_embed / (x_embed[:, -1:] + eps) * scale
    dim_t = torch.arange(num_pos_features, dtype=dtype)
    dim_t = temperature ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / num_pos_features)
    pos_x = x_embed[..., None] / dim_t
    pos_y = y_embed[..., None] / dim_t
    pos_x = torch.stack([pos_x[..., 0::2].sin(), pos_x[..., 1::2].cos()], dim=3).reshape(h, w, -1)
    pos_y = torch.stack([pos_y[..., 0::2].sin(), pos_y[..., 1::2].cos()], dim=3).reshape(h, w, -1)
    return torch.cat([pos_y, pos_x], dim=2)


def linear(x, P, p):
    """custom_layers.py:49-50   x . W^T + b"""
    return x @ P[p + "/kernel"].t() + P[p + "/bias"]


def layer_norm(x, P, p, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), P[p + "/gamma"], P[p + "/beta"], eps)


def mha(P, p, query, key, value, num_heads=8, dropout_p=0.0, gen=None):
    """transformer.py:285-356.  Inputs batch-first [B, L, d] (the reference is sequence-first;
    the arithmetic per (batch, head) is identical).  No masks (they are dead code :322-337)."""
    d = query.shape[-1]
    dh = d // num_heads
    W, b = P[p + "/in_proj_kernel"], P[p + "/in_proj_bias"]
    q = query @ W[:d].t() + b[:d]
    k = key @ W[d:2 * d].t() + b[d:2 * d]
    v = value @ W[2 * d:].t() + b[2 * d:]
    q = q * float(dh) ** -0.5
    B, Lq, _ = q.shape
    Lk = k.shape[1]
    q = q.view(B, Lq, num_heads, dh).transpose(1, 2)
    k = k.view(B, Lk, num_heads, dh).transpose(1, 2)
    v = v.view(B, Lk, num_heads, dh).transpose(1, 2)
    w = torch.softmax(q @ k.transpose(-1, -2), dim=-1)
    if dropout_p > 0:
        keep = (torch.rand(w.shape, generator=gen) >= dropout_p).to(w.dtype)
        w = w * keep / (1 - dropout_p)
    o = (w @ v).transpose(1, 2).reshape(B, Lq, d)
    return o @ P[p + "/out_proj_kernel"].t() + P[p + "/out_proj_bias"]


def _drop(x, p, gen):
    if p <= 0:
        return x
    keep = (torch.rand(x.shape, generator=gen) >= p).to(x.dtype)
    return x * keep / (1 - p)


def encoder_layer(P, p, src, pos, dropout_p=0.0, gen=None):
    """transformer.py:157-179 (post-norm)"""
    q = k = src + pos
    a = mha(P, p + "/self_attn", q, k, src, dropout_p=dropout_p, gen=gen)
    src = layer_norm(src + _drop(a, dropout_p, gen), P, p + "/norm1")
    x = _drop(F.relu(linear(src, P, p + "/linear1")), dropout_p, gen)
    x = linear(x, P, p + "/linear2")
    return layer_norm(src + _drop(x, dropout_p, gen), P, p + "/norm2")


def decoder_layer(P, p, tgt, memory, pos, query_pos, dropout_p=0.0, gen=None):
    """transformer.py:207-234"""
    q = k = tgt + query_pos
    a = mha(P, p + "/self_attn", q, k, tgt, dropout_p=dropout_p, gen=gen)
    tgt = layer_norm(tgt + _drop(a, dropout_p, gen), P, p + "/norm1")
    a = mha(P, p + "/multihead_attn", tgt + query_pos, memory + pos, memory,
            dropout_p=dropout_p, gen=gen)
    tgt = layer_norm(tgt + _drop(a, dropout_p, gen), P, p + "/norm2")
    x = _drop(F.relu(linear(tgt, P, p + "/linear1")), dropout_p, gen)
    x = linear(x, P, p + "/linear2")
    return layer_norm(tgt + _drop(x, dropout_p, gen), P, p + "/norm3")


def transformer_forward(P, src, pos, num_encoder_layers=6, num_decoder_layers=6,
                        dropout_p=0.0, gen=None, return_memory=False):
    """transformer.py:29-57, 74-86, 104-133.  src [B,S,d], pos [S,d] -> hs [L,B,Q,d]."""
    B = src.shape[0]
    pos_b = pos.unsqueeze(0).expand(B, -1, -1)
    query_pos = P["query_embed/kernel"].unsqueeze(0).expand(B, -1, -1)
    x = src
    for l in range(num_encoder_layers):
        x = encoder_layer(P, f"transformer/encoder/layer_{l}", x, pos_b, dropout_p, gen)
    memory = x
    tgt = torch.zeros_like(query_pos)                  # transformer.py:45
    hs = []
    for l in range(num_decoder_layers):
        tgt = decoder_layer(P, f"transformer/decoder/layer_{l}", tgt, memory, pos_b, query_pos,
                            dropout_p, gen)
        hs.append(layer_norm(tgt, P, "transformer/decoder/norm"))   # :122-126
    hs = torch.stack(hs, 0)
    return (hs, memory) if return_memory else hs


def detr_forward(P, images_nhwc, backbone="resnet50", num_encoder_layers=6, num_decoder_layers=6,
                 training=False, dropout_p=0.1, gen=None, return_hs=False):
    """get_detr_model(include_top=True) functional graph, detr.py:141-204."""
    feat = backbone_forward(P, images_nhwc, backbone)                 # [B,h,w,2048]
    B, h, w, C = feat.shape
    proj = feat.reshape(B, h * w, C) @ P["input_proj/kernel"].reshape(C, -1) + P["input_proj/bias"]
    pos = position_embedding_sine(h, w, dtype=feat.dtype).reshape(h * w, -1)
    hs = transformer_forward(P, proj, pos, num_encoder_layers, num_decoder_layers,
                             dropout_p if training else 0.0, gen)
    if "cls_layer/kernel" in P:                         # fine-tuning heads, detr.py:94-114 (Keras Dense: x @ kernel + bias)
        dense = lambda x, p: x @ P[p + "/kernel"] + P[p + "/bias"]
        logits = dense(hs, "cls_layer")
        t = F.relu(dense(hs, "pos_layer/dense"))
        t = F.relu(dense(t, "pos_layer/dense_1"))
        boxes = torch.sigmoid(dense(t, "pos_layer/dense_2"))
    else:
        logits = linear(hs, P, "class_embed")
        t = F.relu(linear(hs, P, "bbox_embed_0"))
        t = F.relu(linear(t, P, "bbox_embed_1"))
        boxes = torch.sigmoid(linear(t, P, "bbox_embed_2"))
    out = {"pred_logits": logits[-1], "pred_boxes": boxes[-1],
           "aux": [{"pred_logits": logits[i], "pred_boxes": boxes[i]}
                   for i in range(num_decoder_layers - 1)]}
    if return_hs:
        out["hs"] = hs
    return out


# --------------------------------------------------------------------------------------
# matcher + set criterion
# --------------------------------------------------------------------------------------

def xcycwh_to_xy_min_xy_max(b):
    """bbox.py:171-183 (clips to [0,1] at :182)"""
    xy = torch.cat([b[:, :2] - b[:, 2:] / 2, b[:, :2] + b[:, 2:] / 2], dim=-1)
    return xy.clamp(0.0, 1.0)


def _pairwise_giou_terms(pa, tb):
    """bbox.py:29-105 (intersect, jaccard) + hungarian_matching.py:186-192. pa[A,4], tb[B,4] xyxy."""
    rb = torch.minimum(pa[:, None, 2:], tb[None, :, 2:])
    lt = torch.maximum(pa[:, None, :2], tb[None, :, :2])
    inter = F.relu(rb - lt)
    inter = inter[..., 0] * inter[..., 1]
    area_a = (pa[:, 2] - pa[:, 0]) * (pa[:, 3] - pa[:, 1])
    area_b = (tb[:, 2] - tb[:, 0]) * (tb[:, 3] - tb[:, 1])
    union = area_a[:, None] + area_b[None, :] - inter
    iou = inter / union
    tl = torch.minimum(pa[:, None, :2], tb[None, :, :2])
    br = torch.maximum(pa[:, None, 2:], tb[None, :, 2:])
    size = F.relu(br - tl)
    area = size[..., 0] * size[..., 1]
    return iou - (area - union) / area


def cost_matrix(t_bbox, t_class, p_bbox, p_class, fcost_class=1, fcost_bbox=5, fcost_giou=2):
    """hungarian_matching.py:165-195.  t_* are in the padded wire format (data/processing.py:35-55):
    row 0 is the header [n,0,0,0].  Returns (C[Q,n], t_bbox[n,4], t_class[n])."""
    n = int(t_bbox[0, 0])
    tb = t_bbox[1:1 + n]
    tc = t_class[1:1 + n].reshape(-1).long()
    p_xy = xcycwh_to_xy_min_xy_max(p_bbox)
    t_xy = xcycwh_to_xy_min_xy_max(tb)
    sm = torch.softmax(p_class, dim=-1)
    cost_class = -sm[:, tc]
    cost_bbox = (p_bbox[:, None, :] - tb[None, :, :]).abs().sum(-1)
    cost_giou = -_pairwise_giou_terms(p_xy, t_xy)
    C = fcost_bbox * cost_bbox + fcost_class * cost_class + fcost_giou * cost_giou
    return C, tb, tc


def hungarian_matching(t_bbox, t_class, p_bbox, p_class, fcost_class=1, fcost_bbox=5, fcost_giou=2):
    """hungarian_matching.py:163-203 + :27-46.  Returns, in the reference's return order,
    (t_indices, p_indices, t_selector, p_selector, t_bbox[n,4], t_class[n]) as seen by the caller
    (loss.py:118): p_indices = ascending query ids (scipy row_ind), t_indices[k] = target matched
    to p_indices[k], p_selector[Q] bool, t_selector[n] bool."""
    from scipy.optimize import linear_sum_assignment
    with torch.no_grad():
        C, tb, tc = cost_matrix(t_bbox, t_class, p_bbox, p_class, fcost_class, fcost_bbox, fcost_giou)
    Cn = C.detach().cpu().numpy()
    rows, cols = linear_sum_assignment(Cn)
    p_sel = np.zeros(Cn.shape[0], dtype=bool)
    p_sel[rows] = True
    t_sel = np.zeros(Cn.shape[1], dtype=bool)
    t_sel[cols] = True
    return (torch.from_numpy(cols.astype(np.int64)), torch.from_numpy(rows.astype(np.int64)),
            torch.from_numpy(t_sel), torch.from_numpy(p_sel), tb, tc)


def get_detr_losses(pred_logits, pred_boxes, t_bbox, t_class, background_class, suffix="",
                    return_indices=False, match_override=None):
    """loss.py:98-179 with loss_labels :37-69 and loss_boxes :72-96 (diag of the NxN GIoU == per-pair).
    match_override [B,Q] (target index per query, -1 = unmatched) replaces the Hungarian step: used by tests to
    compare gradients under an identical assignment (assignments of near-tied random-init predictions flip
    under bf16 noise, which says nothing about the gradient arithmetic)."""
    B, Q, C = pred_logits.shape
    t_idx, p_idx, tb_all, tc_all, p_sel_all = [], [], [], [], []
    t_off = 0
    for b in range(B):
        if match_override is not None:
            mo = match_override[b].long()
            n = int(t_bbox[b][0, 0])
            pi = torch.nonzero(mo >= 0).squeeze(-1)
            ti = mo[pi]
            psel = mo >= 0
            tb, tc = t_bbox[b][1:1 + n], t_class[b][1:1 + n].reshape(-1).long()
        else:
            ti, pi, _, psel, tb, tc = hungarian_matching(t_bbox[b], t_class[b], pred_boxes[b].detach(),
                                                          pred_logits[b].detach())
        t_idx.append(ti + t_off)
        p_idx.append(pi + b * Q)
        tb_all.append(tb)
        tc_all.append(tc)
        p_sel_all.append(psel)
        t_off += tb.shape[0]
    t_idx, p_idx = torch.cat(t_idx), torch.cat(p_idx)
    tb_all, tc_all, p_sel = torch.cat(tb_all), torch.cat(tc_all), torch.cat(p_sel_all)
    logits = pred_logits.reshape(B * Q, C)
    boxes = pred_boxes.reshape(B * Q, 4)
    # ---- loss_labels
    neg_idx = torch.nonzero(~p_sel).squeeze(-1)
    neg_logits = logits[neg_idx]
    pos_logits = logits[p_idx]
    pos_t = tc_all[t_idx]
    neg_t = torch.full((neg_idx.shape[0],), background_class, dtype=torch.long)
    weights = torch.cat([torch.full((neg_idx.shape[0],), 0.1, dtype=logits.dtype),
                         torch.ones(p_idx.shape[0], dtype=logits.dtype)])
    true_neg = (neg_logits.argmax(-1) == background_class).to(logits.dtype).mean()
    true_pos = (pos_logits.argmax(-1) != background_class).to(logits.dtype).mean()
    pos_acc = (pos_logits.argmax(-1) == pos_t).to(logits.dtype).mean()
    ce = F.cross_entropy(torch.cat([neg_logits, pos_logits]), torch.cat([neg_t, pos_t]), reduction="none")
    label_cost = (ce * weights).sum() / weights.sum()
    # ---- loss_boxes
    pb = boxes[p_idx]
    tb = tb_all[t_idx].to(boxes.dtype)
    N = pb.shape[0]
    l1 = (pb - tb).abs().sum() / N
    p_xy, t_xy = xcycwh_to_xy_min_xy_max(pb), xcycwh_to_xy_min_xy_max(tb)
    giou = torch.diagonal(_pairwise_giou_terms(p_xy, t_xy)) if N <= 512 else _diag_giou(p_xy, t_xy)
    giou_loss = (1 - giou).sum() / N
    out = {"label_cost" + suffix: label_cost, "true_neg" + suffix: true_neg,
           "true_pos" + suffix: true_pos, "pos_accuracy" + suffix: pos_acc,
           "giou_loss" + suffix: giou_loss, "l1_loss" + suffix: l1}
    if return_indices:
        return out, (t_idx, p_idx, p_sel)
    return out


def _diag_giou(pa, tb):
    rb = torch.minimum(pa[:, 2:], tb[:, 2:])
    lt = torch.maximum(pa[:, :2], tb[:, :2])
    inter = F.relu(rb - lt)
    inter = inter[:, 0] * inter[:, 1]
    area_a = (pa[:, 2] - pa[:, 0]) * (pa[:, 3] - pa[:, 1])
    area_b = (tb[:, 2] - tb[:, 0]) * (tb[:, 3] - tb[:, 1])
    union = area_a + area_b - inter
    iou = inter / union
    tl = torch.minimum(pa[:, :2], tb[:, :2])
    br = torch.maximum(pa[:, 2:], tb[:, 2:])
    size = F.relu(br - tl)
    area = size[:, 0] * size[:, 1]
    return iou - (area - union) / area


def get_total_loss(losses):
    """loss.py:6-19 (substring match; aux layers not down-weighted)"""
    total = 0
    for k, v in losses.items():
        for name, w in (("label_cost", 1), ("giou_loss", 2), ("l1_loss", 5)):
            if name in k:
                total = total + v * w
    return total


def get_losses(m_outputs, t_bbox, t_class, background_class, match_override=None):
    """loss.py:22-34.  match_override: optional [L,B,Q] tensor (layer order 0..L-1, main output = last)."""
    naux = len(m_outputs.get("aux", []))
    losses = get_detr_losses(m_outputs["pred_logits"], m_outputs["pred_boxes"], t_bbox, t_class,
                             background_class, match_override=None if match_override is None else match_override[naux])
    for a, aux in enumerate(m_outputs.get("aux", [])):
        losses.update(get_detr_losses(aux["pred_logits"], aux["pred_boxes"], t_bbox, t_class,
                                      background_class, suffix=f"_{a}",
                                      match_override=None if match_override is None else match_override[a]))
    return get_total_loss(losses), losses


# --------------------------------------------------------------------------------------
# targets wire format + synthetic data
# --------------------------------------------------------------------------------------

def pad_labels(boxes, classes):
    """data/processing.py:35-55: boxes [n,4] cxcywh in [0,1], classes [n] -> ([100,4] f32, [100,1] i64)."""
    n = boxes.shape[0]
    assert n <= 99
    tb = torch.zeros(100, 4, dtype=torch.float32)
    tc = torch.zeros(100, 1, dtype=torch.int64)
    tb[0, 0] = float(n)
    tb[1:1 + n] = boxes
    tc[1:1 + n, 0] = classes
    return tb, tc


def normalized_images(image_u8, method="torch_resnet"):
    """data/processing.py:6-23 (numpy float64 arithmetic, then float32): image [...,3] uint8 (or any numeric) pixels 0..255.
    torch_resnet: (x/255 - mean)/std per RGB channel; tf_resnet: RGB -> BGR, minus the caffe means."""
    image = np.asarray(image_u8)
    if method == "torch_resnet":
        channel_avg = np.array([0.485, 0.456, 0.406])
        channel_std = np.array([0.229, 0.224, 0.225])
        return ((image / 255.0 - channel_avg) / channel_std).astype(np.float32)
    if method == "tf_resnet":
        return (image[..., ::-1] - np.array([103.939, 116.779, 123.68])).astype(np.float32)
    raise ValueError(method)


def get_model_inference(m_outputs, background_class, bbox_format="xy_center"):
    """inference.py:68-95 -- image 0 of the batch: softmax -> (max score, argmax label, first index on ties) -> drop the
    queries whose label is the background class (ascending query order) -> boxes as cxcywh | clipped xyxy | clipped yxyx."""
    boxes = torch.as_tensor(m_outputs["pred_boxes"])[0].float()
    logits = torch.as_tensor(m_outputs["pred_logits"])[0].float()
    sm = torch.softmax(logits, -1)
    scores, labels = sm.max(-1)
    labels = torch.from_numpy(np.argmax(sm.numpy(), -1))         # numpy argmax = first index, like tf.argmax
    keep = torch.nonzero(labels != background_class).squeeze(-1)
    scores, labels, boxes = scores[keep], labels[keep], boxes[keep]
    if bbox_format == "xyxy":
        boxes = xcycwh_to_xy_min_xy_max(boxes)
    elif bbox_format == "yxyx":
        boxes = xcycwh_to_xy_min_xy_max(boxes)[:, [1, 0, 3, 2]]
    elif bbox_format != "xy_center":
        raise NotImplementedError()
    return boxes, labels, scores


def synthetic_targets(B, n=20, num_classes=91, seed=0, n_range=None):
    """cx,cy~U(.1,.9), w,h~U(.02,.5), class~randint(0,91)."""
    g = torch.Generator().manual_seed(seed)
    tbs, tcs = [], []
    for b in range(B):
        nb = n if n_range is None else int(torch.randint(n_range[0], n_range[1] + 1, (1,), generator=g))
        cxcy = torch.rand(nb, 2, generator=g) * 0.8 + 0.1
        wh = torch.rand(nb, 2, generator=g) * 0.48 + 0.02
        cls = torch.randint(0, num_classes, (nb,), generator=g)
        tb, tc = pad_labels(torch.cat([cxcy, wh], 1), cls)
        tbs.append(tb)
        tcs.append(tc)
    return torch.stack(tbs), torch.stack(tcs)


# --------------------------------------------------------------------------------------
# optimizer glue  (optimizers.py)
# --------------------------------------------------------------------------------------

def param_group(name):
    """optimizers.py:10-43 for the include_top=True model: backbone = ResNet convs + input_proj;
    transformers = transformer.* + class_embed + bbox_embed_*; query_embed in NO group;
    BN vectors are non-trainable (custom_layers.py:11-18)."""
    if "/bn" in name or "downsample_1" in name:
        return None
    if name.startswith("backbone/") or name.startswith("input_proj/"):
        return "backbone"
    if name.startswith("query_embed"):
        return None
    if name.startswith("cls_layer/") or name.startswith("pos_layer/"):
        return "nlayers"                # optimizers.py:39-43 (config.nlayers = ["cls_layer", "pos_layer"], detr.py:103)
    return "transformers"


def adam_clipnorm_step(param, grad, m, v, step, lr, clipnorm=0.1, beta1=0.9, beta2=0.999, eps=1e-7):
    """Keras (TF 2.3) Adam.apply_gradients with clipnorm: each gradient tensor is clipped to L2
    norm `clipnorm` on its own (tf.clip_by_norm), then the standard Keras Adam update
    lr_t = lr*sqrt(1-b2^t)/(1-b1^t); p -= lr_t * m / (sqrt(v) + eps)   (optimizers.py:86-88)."""
    norm = grad.norm()
    if norm > clipnorm:
        grad = grad * (clipnorm / norm)
    m.mul_(beta1).add_(grad, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(grad, grad, value=1 - beta2)
    lr_t = lr * math.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
    param.sub_(lr_t * m / (v.sqrt() + eps))
    return param


def train_step(P, images, t_bbox, t_class, background_class=91, backbone="resnet50",
               num_encoder_layers=6, num_decoder_layers=6, gradient_aggregate=1, training=False,
               match_override=None):
    """training.py:9-25: fwd -> get_losses -> /gradient_aggregate -> grads for every trainable var."""
    names = [n for n in P if param_group(n) is not None]
    Pg = OrderedDict((n, (p.clone().requires_grad_(True) if n in names else p)) for n, p in P.items())
    out = detr_forward(Pg, images, backbone, num_encoder_layers, num_decoder_layers, training=training)
    total, log = get_losses(out, t_bbox, t_class, background_class, match_override)
    total = total / gradient_aggregate
    grads = torch.autograd.grad(total, [Pg[n] for n in names], allow_unused=True)
    return out, total.detach(), {k: v.detach() for k, v in log.items()}, dict(zip(names, grads))
