"""CPU restatement of the reference's mAP evaluation -- TEST INFRASTRUCTURE (the checker), never the product path.

Follows /root/reference/detr_tf/loss/compute_map.py: APDataObject.get_ap (:35-81), compute_iou / compute_overlaps (:104-139),
cal_map (:183-272, the 'box' entries; the reference's eval.py:51 feeds zero masks, whose IoU is 0/0 = NaN: no 'mask' detection
ever matches), calc_map (:142-171); driven as in eval.py:30-61.  Pinned against tests/golden/map_golden.npz, which
tests/golden/make_golden_map.py produces by executing the reference's own compute_map.py.
"""
from collections import OrderedDict

import numpy as np


def yxyx_from_xcycwh(b):
    """bbox.py:171-183 (corners clipped to [0,1]) + :125-138 (yx order), float32"""
    b = np.asarray(b, np.float32).reshape(-1, 4)
    xy = np.concatenate([b[:, :2] - b[:, 2:] / np.float32(2), b[:, :2] + b[:, 2:] / np.float32(2)], -1)
    xy = np.clip(xy, np.float32(0), np.float32(1))
    return xy[:, [1, 0, 3, 2]]


def overlaps(boxes1, boxes2):
    """compute_map.py:124-139: IoU[boxes1, boxes2] in float32 arithmetic, stored as float64"""
    boxes1, boxes2 = np.asarray(boxes1, np.float32).reshape(-1, 4), np.asarray(boxes2, np.float32).reshape(-1, 4)
    area1 = (boxes1[:, 2] - boxes1[:, 0]) * (boxes1[:, 3] - boxes1[:, 1])
    area2 = (boxes2[:, 2] - boxes2[:, 0]) * (boxes2[:, 3] - boxes2[:, 1])
    out = np.zeros((boxes1.shape[0], boxes2.shape[0]))
    with np.errstate(invalid="ignore", divide="ignore"):
        for i in range(out.shape[1]):
            box = boxes2[i]
            y1, y2 = np.maximum(box[0], boxes1[:, 0]), np.minimum(box[2], boxes1[:, 2])
            x1, x2 = np.maximum(box[1], boxes1[:, 1]), np.minimum(box[3], boxes1[:, 3])
            inter = np.maximum(x2 - x1, 0) * np.maximum(y2 - y1, 0)
            union = area2[i] + area1 - inter
            out[:, i] = inter / union
    return out


class APData:
    """compute_map.py:16-81"""

    def __init__(self):
        self.data_points, self.num_gt_positives = [], 0

    def is_empty(self):
        return len(self.data_points) == 0 and self.num_gt_positives == 0

    def get_ap(self):
        if self.num_gt_positives == 0:
            return 0
        pts = sorted(self.data_points, key=lambda x: -x[0])            # stable
        precisions, recalls, nt, nf = [], [], 0, 0
        for _, ok in pts:
            if ok:
                nt += 1
            else:
                nf += 1
            precisions.append(nt / (nt + nf))
            recalls.append(nt / self.num_gt_positives)
        for i in range(len(precisions) - 1, 0, -1):
            if precisions[i] > precisions[i - 1]:
                precisions[i - 1] = precisions[i]
        y = [0] * 101
        idx = np.searchsorted(np.array(recalls), np.array([x / 100 for x in range(101)]), side="left")
        for bar, pi in enumerate(idx):
            if pi < len(precisions):
                y[bar] = precisions[pi]
        return sum(y) / len(y)


def cal_map_image(p_bbox, p_labels, p_scores, t_bbox, t_classes, ap_box, iou_thresholds):
    """compute_map.py:183-272 for one image, 'box' entries; boxes in yxyx.  ap_box[iou_idx][class] are APData."""
    classes, scores = [int(c) for c in p_labels], [float(s) for s in p_scores]
    gt = [int(c) for c in t_classes]
    iou = overlaps(p_bbox, t_bbox)
    order = sorted(range(len(classes)), key=lambda i: -scores[i])
    for cls in set(classes + gt):
        ngt = sum(1 for x in gt if x == cls)
        for a, thr in enumerate(iou_thresholds):
            used = [False] * len(gt)
            obj = ap_box[a][cls]
            obj.num_gt_positives += ngt
            for i in order:
                if classes[i] != cls:
                    continue
                best, bj = thr, -1
                for j in range(len(gt)):
                    if used[j] or gt[j] != cls:
                        continue
                    v = iou[i, j].item()
                    if v > best:
                        best, bj = v, j
                if bj >= 0:
                    used[bj] = True
                    obj.data_points.append((scores[i], True))
                else:
                    obj.data_points.append((scores[i], False))


def calc_map(ap_box, iou_thresholds, num_classes):
    """compute_map.py:142-171 ('box' row): {'all': ., 50: ., ... 95: .} rounded to two decimals"""
    out = OrderedDict()
    out["all"] = 0
    for a, thr in enumerate(iou_thresholds):
        aps = [ap_box[a][c].get_ap() for c in range(num_classes) if not ap_box[a][c].is_empty()]
        out[int(thr * 100)] = sum(aps) / len(aps) * 100 if aps else 0
    out["all"] = sum(out.values()) / (len(out) - 1)
    return OrderedDict((k, round(v, 2)) for k, v in out.items())
