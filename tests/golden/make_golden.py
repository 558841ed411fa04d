"""Generate golden vectors for the loss / matcher path by EXECUTING THE REFERENCE'S OWN CODE.

    python tests/golden/make_golden.py            # needs /root/reference (build container only)

detr_tf/loss/loss.py, detr_tf/loss/hungarian_matching.py and detr_tf/bbox.py are imported
unmodified from /root/reference with `tensorflow` replaced by tests/golden/tf_numpy_shim.py
(TF itself is not installable here) and with the real scipy.optimize.linear_sum_assignment.
Outputs -> tests/golden/loss_golden.npz (committed; /root/reference does not exist on the GPU box).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.path.insert(0, "/root/reference")
from detr_tf.loss import hungarian_matching as ref_hm  # noqa: E402
from detr_tf.loss import loss as ref_loss  # noqa: E402


class Cfg:
    background_class = 91


def make_targets(rs, B, ns):
    t_bbox = np.zeros((B, 100, 4), np.float32)
    t_class = np.zeros((B, 100, 1), np.int64)
    for b in range(B):
        n = ns[b]
        t_bbox[b, 0, 0] = n
        t_bbox[b, 1:1 + n, :2] = rs.uniform(0.1, 0.9, (n, 2))
        t_bbox[b, 1:1 + n, 2:] = rs.uniform(0.02, 0.5, (n, 2))
        t_class[b, 1:1 + n, 0] = rs.randint(0, 91, n)
    return t_bbox, t_class


def make_preds(rs, L, B, Q=100, C=92):
    logits = rs.randn(L, B, Q, C).astype(np.float32)
    boxes = np.concatenate([rs.uniform(0.05, 0.95, (L, B, Q, 2)), rs.uniform(0.02, 0.5, (L, B, Q, 2))],
                           -1).astype(np.float32)
    # a few boxes hanging over the image border so the [0,1] clip (bbox.py:182) is exercised
    boxes[:, :, ::7, 2:] *= 2.5
    return logits, boxes


def main():
    out = {}
    rs = np.random.RandomState(1234)

    # ---- matcher cases: cost matrix + indices, per image
    captured = []
    orig = ref_hm.np_tf_linear_sum_assignment

    def spy(matrix):
        captured.append(np.array(matrix, copy=True))
        return orig(matrix)
    ref_hm.np_tf_linear_sum_assignment = spy

    ns = [20, 1, 7, 99, 50, 3]
    B = len(ns)
    t_bbox, t_class = make_targets(rs, B, ns)
    logits, boxes = make_preds(rs, 1, B)
    out["m_t_bbox"], out["m_t_class"] = t_bbox, t_class
    out["m_logits"], out["m_boxes"] = logits[0], boxes[0]
    for b in range(B):
        t_idx, p_idx, t_sel, p_sel, tb, tc = ref_hm.hungarian_matching(
            t_bbox[b], t_class[b], boxes[0, b], logits[0, b], slice_preds=True)
        out[f"m_cost_{b}"] = captured[-1].astype(np.float32)
        out[f"m_t_indices_{b}"] = np.asarray(t_idx, np.int64)
        out[f"m_p_indices_{b}"] = np.asarray(p_idx, np.int64)
        out[f"m_t_selector_{b}"] = np.asarray(t_sel, bool)
        out[f"m_p_selector_{b}"] = np.asarray(p_sel, bool)
        out[f"m_tb_{b}"] = np.asarray(tb, np.float32)
        out[f"m_tc_{b}"] = np.asarray(tc, np.int64)

    # ---- full set criterion over 6 decoder layers (loss.py:22-34)
    ns = [20, 5, 11]
    B = len(ns)
    t_bbox, t_class = make_targets(rs, B, ns)
    logits, boxes = make_preds(rs, 6, B)
    m_outputs = {"pred_logits": logits[5], "pred_boxes": boxes[5],
                 "aux": [{"pred_logits": logits[i], "pred_boxes": boxes[i]} for i in range(5)]}
    total, losses = ref_loss.get_losses(m_outputs, t_bbox, t_class, Cfg())
    out["l_t_bbox"], out["l_t_class"] = t_bbox, t_class
    out["l_logits"], out["l_boxes"] = logits, boxes
    out["l_total"] = np.float32(total)
    keys = sorted(losses.keys())
    out["l_keys"] = np.array(keys)
    out["l_values"] = np.array([np.float32(losses[k]) for k in keys], np.float32)

    path = os.path.join(HERE, "loss_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; total loss", float(total))
    for k in keys:
        if not k[-1].isdigit():
            print(" ", k, float(losses[k]))


if __name__ == "__main__":
    main()
