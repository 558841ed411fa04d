"""get_losses -- mirror of detr_tf/loss/loss.py:22-34 on device tensors (matcher + set criterion kernels)."""
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib, ops

NAMES = ("label_cost", "true_neg", "true_pos", "pos_accuracy", "giou_loss", "l1_loss")


def _dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype).contiguous()


def get_losses(m_outputs, t_bbox, t_class, config):
    """m_outputs: {'pred_logits': [B,Q,C], 'pred_boxes': [B,Q,4], 'aux': [...]}; t_bbox [B,100,4], t_class [B,100,1]
    (wire format data/processing.py:35-55).  Returns (total_loss, losses) with the reference's 36 keys; every value is a
    0-dim device tensor (no host sync).  A cost matrix with NaN / -inf entries -- the reference raises through scipy,
    hungarian_matching.py:29 -- makes every returned value NaN (checked on device, visible at the caller's read-back)."""
    layers = list(m_outputs.get("aux", [])) + [m_outputs]
    device = layers[0]["pred_logits"].device if isinstance(layers[0]["pred_logits"], torch.Tensor) else torch.device("cuda")
    if device.type != "cuda" and not getattr(_lib, "_EMULATED", False):
        device = torch.device("cuda")            # host tensors (numpy / CPU torch) are copied in: the kernels take device pointers only
    logits = torch.stack([_dev(l["pred_logits"], torch.float32, device) for l in layers])      # [L,B,Q,C]
    boxes = torch.stack([_dev(l["pred_boxes"], torch.float32, device) for l in layers])
    L, B, Q, C = logits.shape
    tb = _dev(t_bbox, torch.float32, device).reshape(B, 100, 4)
    tc = _dev(t_class, torch.int64, device).reshape(B, 100, 1)
    P = L * B
    p_idx = torch.empty(P, Q, dtype=torch.int64, device=device)
    t_idx = torch.empty(P, Q, dtype=torch.int64, device=device)
    p_sel = torch.empty(P, Q, dtype=torch.uint8, device=device)
    match = torch.empty(P, Q, dtype=torch.int32, device=device)
    status = torch.empty(P, dtype=torch.int32, device=device)
    ops.matcher(logits, C, boxes, tb, tc, P, B, Q, C, p_idx, t_idx, p_sel, match, None, status)
    sums = torch.empty(L, 8, dtype=torch.float32, device=device)
    losses = torch.empty(L, 6, dtype=torch.float32, device=device)
    total = torch.empty(1, dtype=torch.float32, device=device)
    ops.set_loss(logits, C, boxes, tb, tc, match, L, B, Q, C, int(config.background_class), None, 1.0, sums, losses,
                 total, None, 0, None, 0, status=status)
    out = OrderedDict()
    for l in [L - 1] + list(range(L - 1)):
        suf = "" if l == L - 1 else f"_{l}"
        for k, n in enumerate(NAMES):
            out[n + suf] = losses[l, k]
    return total[0], out
