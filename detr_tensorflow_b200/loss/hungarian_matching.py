"""hungarian_matching -- mirror of detr_tf/loss/hungarian_matching.py:163-203, computed on device
(cost matrix + exact assignment in csrc/matcher.cu; no host round trip, no scipy)."""
import numpy as np
import torch

from .. import ops


def _dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype).contiguous()


def hungarian_matching(t_bbox, t_class, p_bbox, p_class, fcost_class=1, fcost_bbox=5, fcost_giou=2, slice_preds=True,
                       device="cuda"):
    """One image.  t_bbox [100,4] / t_class [100,1] in the padded wire format (data/processing.py:35-55) when
    slice_preds=True, else already sliced [n,4] / [n].  Returns the reference's 6-tuple in the reference's order, as
    seen by its caller (loss.py:118): (t_indices, p_indices, t_selector, p_selector, t_bbox[n,4], t_class[n])."""
    device = torch.device(device)
    p_bbox = _dev(p_bbox, torch.float32, device)
    p_class = _dev(p_class, torch.float32, device)
    Q, C = p_class.shape
    if slice_preds:
        tb = _dev(t_bbox, torch.float32, device).reshape(1, 100, 4)
        tc = _dev(t_class, torch.int64, device).reshape(1, 100, 1)
    else:
        tbs = _dev(t_bbox, torch.float32, device)
        n = tbs.shape[0]
        tb = torch.zeros(1, 100, 4, dtype=torch.float32, device=device)
        tc = torch.zeros(1, 100, 1, dtype=torch.int64, device=device)
        tb[0, 0, 0] = n
        tb[0, 1:1 + n] = tbs
        tc[0, 1:1 + n, 0] = _dev(t_class, torch.int64, device).reshape(-1)
    p_idx = torch.empty(1, Q, dtype=torch.int64, device=device)
    t_idx = torch.empty(1, Q, dtype=torch.int64, device=device)
    p_sel = torch.empty(1, Q, dtype=torch.uint8, device=device)
    match = torch.empty(1, Q, dtype=torch.int32, device=device)
    status = torch.empty(1, dtype=torch.int32, device=device)
    ops.matcher(p_class, C, p_bbox, tb, tc, 1, 1, Q, C, p_idx, t_idx, p_sel, match, None, status,
                float(fcost_class), float(fcost_bbox), float(fcost_giou))
    n = int(tb[0, 0, 0])                      # the reference's return shapes depend on n (host value needed)
    if int(status[0]) != 0:
        raise ValueError("matrix contains invalid numeric entries")   # scipy's error for NaN / -inf costs
    t_selector = torch.ones(n, dtype=torch.bool, device=device)
    return (t_idx[0, :n], p_idx[0, :n], t_selector, p_sel[0].bool(), tb[0, 1:1 + n], tc[0, 1:1 + n, 0])
