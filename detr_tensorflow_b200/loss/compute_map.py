"""Mirror of detr_tf/loss/compute_map.py (APDataObject :16-81, cal_map :183-272, calc_map :142-171, print_maps :173-181) and of the
evaluation loop of the reference's eval.py:30-61, with the per-image matching on device (csrc/pipeline.cu:map_match_kernel: stable
descending-score order, greedy assignment per class and IoU threshold) for a whole batch per launch.

Masks: the reference's eval.py:51 feeds all-zero masks to cal_map, whose mask IoU is 0/0 = NaN, so no 'mask' detection ever
matches; the 'mask' entries are kept for format compatibility and filled the same way (every detection a false positive).
The precision/recall integration (get_ap) runs once per evaluation on the host, vectorised, in float64 like the reference."""
from collections import OrderedDict

import numpy as np
import torch

from .. import _lib, ops
from ..inference import batched_model_inference

IOU_THRESHOLDS = [x / 100. for x in range(50, 100, 5)]            # eval.py:34


class APDataObject:
    """compute_map.py:16-81: the (score, is_true) points of one class at one IoU threshold"""

    def __init__(self):
        self.data_points = []
        self.num_gt_positives = 0

    def push(self, score, is_true):
        self.data_points.append((score, is_true))

    def extend(self, scores, flags):
        self.data_points.extend(zip(scores, flags))

    def add_gt_positives(self, num_positives):
        self.num_gt_positives += num_positives

    def is_empty(self):
        return len(self.data_points) == 0 and self.num_gt_positives == 0

    def get_ap(self):
        """:35-81: sort by score (stable), running precision / recall, right-to-left precision envelope, 101-point average"""
        if self.num_gt_positives == 0:
            return 0
        if not self.data_points:
            return 0.0
        scores = np.array([p[0] for p in self.data_points], np.float64)
        flags = np.array([bool(p[1]) for p in self.data_points])
        order = np.argsort(-scores, kind="stable")
        tp = np.cumsum(flags[order])
        k = np.arange(1, len(tp) + 1)
        precisions = tp / k
        recalls = tp / self.num_gt_positives
        precisions = np.maximum.accumulate(precisions[::-1])[::-1]
        idx = np.searchsorted(recalls, np.array([x / 100 for x in range(101)]), side="left")
        y = np.where(idx < len(precisions), precisions[np.minimum(idx, len(precisions) - 1)], 0.0)
        return sum(y.tolist()) / 101                       # (python summation order, like the reference's sum(y_range))


def _dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return torch.as_tensor(x).to(device=device, dtype=dtype).contiguous()


def _device(x):
    if isinstance(x, torch.Tensor) and x.is_cuda:
        return x.device
    return torch.device("cpu") if getattr(_lib, "_EMULATED", False) else torch.device("cuda")


def _match(pred_boxes, pred_labels, pred_scores, pred_count, t_boxes, t_labels, t_count, t_wire, iou_thresholds, num_classes):
    """one launch of map_match_kernel -> (rank [B,Q] i32, tp [B,T,Q] u8, gt_count [num_classes] i32), device tensors"""
    B, Q = pred_labels.shape
    dev = pred_boxes.device
    T = len(iou_thresholds)
    thr = torch.tensor(list(iou_thresholds), dtype=torch.float64).to(dev)
    rank = torch.empty(B, Q, dtype=torch.int32, device=dev)
    tp = torch.empty(B, T, Q, dtype=torch.uint8, device=dev)
    gt_count = torch.zeros(num_classes, dtype=torch.int32, device=dev)
    NT = t_boxes.shape[1]
    ops.map_match(pred_boxes, pred_labels, pred_scores, pred_count, B, Q, t_boxes, t_labels, t_count, NT, t_wire, thr, T, num_classes,
                  rank, tp, gt_count)
    return rank, tp, gt_count


def _push(ap_data, iou_thresholds, labels, scores, rank, tp, count, gt_classes):
    """append one image's detections to the per-(threshold, class) accumulators in the reference's order (class by class,
    descending score within the class)"""
    k = int(count)
    order = np.argsort(rank[:k], kind="stable")               # detections in descending-score order
    lab, sc = labels[:k][order], scores[:k][order]
    for cls in sorted(set(int(c) for c in lab) | set(int(c) for c in gt_classes)):
        sel = lab == cls
        ngt = int(np.sum(np.asarray(gt_classes) == cls))
        for a in range(len(iou_thresholds)):
            obj = ap_data["box"][a][cls]
            obj.add_gt_positives(ngt)
            obj.extend(sc[sel].tolist(), tp[a, :k][order][sel].astype(bool).tolist())
            if "mask" in ap_data:                             # zero masks: IoU = NaN, nothing ever matches
                m = ap_data["mask"][a][cls]
                m.add_gt_positives(ngt)
                m.extend(sc[sel].tolist(), [False] * int(sel.sum()))


def cal_map(p_bbox, p_labels, p_scores, p_mask, t_bbox, gt_classes, t_mask, ap_data, iou_thresholds):
    """compute_map.py:183-272, one image: p_bbox [k,4] / t_bbox [n,4] in yxyx corners (eval.py:40-49), labels, scores; the masks
    are ignored (see the module docstring).  Pushes into ap_data['box'][iou_idx][class] (and 'mask')."""
    dev = _device(p_bbox)
    pb = _dev(p_bbox, torch.float32, dev).reshape(-1, 4)
    k = pb.shape[0]
    Q = max(k, 1)
    tb = _dev(t_bbox, torch.float32, dev).reshape(-1, 4)
    n = tb.shape[0]
    if Q > 256 or n > 100:
        raise ValueError("cal_map: at most 256 detections and 100 ground-truth boxes per image")
    boxes = torch.zeros(1, Q, 4, dtype=torch.float32, device=dev)
    labels = torch.zeros(1, Q, dtype=torch.int64, device=dev)
    scores = torch.zeros(1, Q, dtype=torch.float32, device=dev)
    boxes[0, :k], labels[0, :k], scores[0, :k] = pb, _dev(p_labels, torch.int64, dev).reshape(-1), _dev(p_scores, torch.float32, dev).reshape(-1)
    tbx = torch.zeros(1, max(n, 1), 4, dtype=torch.float32, device=dev)
    tlb = torch.zeros(1, max(n, 1), dtype=torch.int64, device=dev)
    gt = np.asarray(torch.as_tensor(gt_classes).cpu()).reshape(-1).astype(np.int64)
    tbx[0, :n], tlb[0, :n] = tb, torch.from_numpy(gt).to(dev)
    cnt = torch.tensor([k], dtype=torch.int32).to(dev)
    tcnt = torch.tensor([n], dtype=torch.int32).to(dev)
    ncls = len(ap_data["box"][0])
    rank, tp, _ = _match(boxes, labels, scores, cnt, tbx, tlb, tcnt, False, iou_thresholds, ncls)
    _push(ap_data, iou_thresholds, labels[0].cpu().numpy(), scores[0].cpu().numpy().astype(np.float64), rank[0].cpu().numpy(),
          tp[0].cpu().numpy(), k, gt)


def calc_map(ap_data, iou_thresholds, class_name, print_result=False):
    """compute_map.py:142-171: {'box': {'all', 50, 55, ... 95}, 'mask': {...}} in percent, rounded to two decimals"""
    types = [t for t in ("box", "mask") if t in ap_data]
    aps = [{t: [] for t in types} for _ in iou_thresholds]
    for _class in range(len(class_name)):
        for iou_idx in range(len(iou_thresholds)):
            for iou_type in types:
                ap_obj = ap_data[iou_type][iou_idx][_class]
                if not ap_obj.is_empty():
                    aps[iou_idx][iou_type].append(ap_obj.get_ap())
    all_maps = {t: OrderedDict() for t in types}
    for iou_type in types:
        all_maps[iou_type]["all"] = 0
        for i, threshold in enumerate(iou_thresholds):
            vals = aps[i][iou_type]
            all_maps[iou_type][int(threshold * 100)] = sum(vals) / len(vals) * 100 if len(vals) > 0 else 0
        all_maps[iou_type]["all"] = sum(all_maps[iou_type].values()) / (len(all_maps[iou_type].values()) - 1)
    if print_result:
        print_maps(all_maps)
    return {k: {j: round(u, 2) for j, u in v.items()} for k, v in all_maps.items()}


def print_maps(all_maps):
    """compute_map.py:173-181"""
    first = next(iter(all_maps.values()))
    make_row = lambda vals: (" %5s |" * len(vals)) % tuple(vals)
    make_sep = lambda n: ("-------+" * n)
    print()
    print(make_row([""] + [(".%d " % x if isinstance(x, int) else x + " ") for x in first.keys()]))
    print(make_sep(len(first) + 1))
    for iou_type, row in all_maps.items():
        print(make_row([iou_type] + ["%.2f" % x if x < 100 else "%.1f" % x for x in row.values()]))
    print(make_sep(len(first) + 1))
    print()


class MapEvaluator:
    """Batched evaluation (extension of eval.py:30-61, which runs one image per step): per batch ONE post-process launch and ONE
    matching launch on device; the small per-detection results (score, label, rank, true-positive flags) are kept on the device
    until summary(), which copies them to the host once and integrates the precision/recall curves."""

    def __init__(self, class_names, iou_thresholds=None, with_mask_rows=True):
        self.class_names = list(class_names)
        self.iou_thresholds = list(iou_thresholds) if iou_thresholds is not None else list(IOU_THRESHOLDS)
        self.with_mask_rows = with_mask_rows
        self.batches = []

    def update(self, m_outputs, t_bbox, t_class, background_class):
        """m_outputs: the model's output dict for a batch; t_bbox [B,100,4] / t_class [B,100,1]: the padded wire format"""
        boxes, labels, scores, _, count = batched_model_inference(m_outputs, background_class, bbox_format="yxyx")
        dev = boxes.device
        tb = _dev(t_bbox, torch.float32, dev).reshape(-1, 100, 4)
        tc = _dev(t_class, torch.int64, dev).reshape(-1, 100, 1)
        rank, tp, _ = _match(boxes, labels, scores, count, tb, tc, None, True, self.iou_thresholds, len(self.class_names))
        self.batches.append((labels, scores, rank, tp, count, tb[:, 0, 0].clone(), tc[:, :, 0].clone()))

    def ap_data(self):
        T, C = len(self.iou_thresholds), len(self.class_names)
        ap = {"box": [[APDataObject() for _ in range(C)] for _ in range(T)]}
        if self.with_mask_rows:
            ap["mask"] = [[APDataObject() for _ in range(C)] for _ in range(T)]
        for labels, scores, rank, tp, count, tn, tcls in self.batches:
            labels, scores, rank, tp = labels.cpu().numpy(), scores.cpu().numpy().astype(np.float64), rank.cpu().numpy(), tp.cpu().numpy()
            count, tn, tcls = count.cpu().numpy(), tn.cpu().numpy(), tcls.cpu().numpy()
            for b in range(labels.shape[0]):
                n = int(min(max(tn[b], 0), 99))
                _push(ap, self.iou_thresholds, labels[b], scores[b], rank[b], tp[b], count[b], tcls[b, 1:1 + n])
        return ap

    def summary(self, print_result=False):
        return calc_map(self.ap_data(), self.iou_thresholds, self.class_names, print_result=print_result)


def eval_model(model, config, class_names, valid_dt, print_result=True):
    """eval.py:30-61: forward (training=False) -> get_model_inference(yxyx) -> cal_map per image -> calc_map; batched on device.
    valid_dt: iterable of (images, t_bbox [B,100,4], t_class [B,100,1]).  Returns the summary dict of calc_map."""
    ev = MapEvaluator(class_names)
    for it, (images, t_bbox, t_class) in enumerate(valid_dt):
        m_outputs = model(images, training=False)
        ev.update(m_outputs, t_bbox, t_class, config.background_class)
        print(f"Computing map.....{it}", end="\r")
    return ev.summary(print_result=print_result)
