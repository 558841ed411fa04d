"""Golden vectors for the evaluation row (mAP), by EXECUTING THE REFERENCE'S OWN CODE:

    python tests/golden/make_golden_map.py      # needs /root/reference (build container only)

detr_tf/loss/compute_map.py (APDataObject, cal_map, calc_map: :16-81, :183-272, :142-171) is imported unmodified from
/root/reference (tensorflow -> the numpy shim; cv2 / matplotlib stubbed: only plotting helpers use them) and driven the way the
reference's eval.py:30-61 drives it: per image the post-processed predictions in yxyx format, the unpadded targets in yxyx
format, zero masks.  Output -> tests/golden/map_golden.npz (committed).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
for name in ("cv2", "matplotlib", "matplotlib.pyplot"):
    sys.modules.setdefault(name, types.ModuleType(name))
sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
sys.path.insert(0, "/root/reference")
from detr_tf.loss import compute_map as ref_map  # noqa: E402


def yxyx(b):
    """bbox.py:171-183 + :125-138 (xcycwh -> clipped corners -> yx order), float32 like the TF ops"""
    b = b.astype(np.float32)
    xy = np.concatenate([b[:, :2] - b[:, 2:] / np.float32(2), b[:, :2] + b[:, 2:] / np.float32(2)], -1)
    xy = np.clip(xy, np.float32(0), np.float32(1))
    return xy[:, [1, 0, 3, 2]]


def main():
    rs = np.random.RandomState(77)
    ncls, nimg = 6, 12
    class_names = [f"c{i}" for i in range(ncls)]
    iou_thresholds = [x / 100. for x in range(50, 100, 5)]
    ap_data = {"box": [[ref_map.APDataObject() for _ in class_names] for _ in iou_thresholds],
               "mask": [[ref_map.APDataObject() for _ in class_names] for _ in iou_thresholds]}
    out = {"ncls": np.int64(ncls), "nimg": np.int64(nimg)}
    for i in range(nimg):
        n = int(rs.randint(1, 9))
        t = np.concatenate([rs.uniform(0.15, 0.85, (n, 2)), rs.uniform(0.05, 0.4, (n, 2))], -1).astype(np.float32)
        tcls = rs.randint(0, ncls - 1, n).astype(np.int64)                  # the last class never appears as ground truth
        k = int(rs.randint(0, 14))
        # predictions: jittered copies of some targets (true positives at some thresholds), duplicates, and random boxes
        src = rs.randint(0, n, k)
        p = t[src] + rs.normal(0, 0.03, (k, 4)).astype(np.float32) * (rs.rand(k, 1) < 0.8)
        rnd = rs.rand(k) < 0.25
        p[rnd] = np.concatenate([rs.uniform(0.1, 0.9, (rnd.sum(), 2)), rs.uniform(0.05, 0.5, (rnd.sum(), 2))], -1)
        p = p.astype(np.float32)
        pcls = np.where(rs.rand(k) < 0.8, tcls[src], rs.randint(0, ncls, k)).astype(np.int64)
        score = rs.uniform(0.05, 1.0, k).astype(np.float32)
        if k >= 4:
            score[1] = score[0]                                             # score ties: the reference's sorts are stable
            score[3] = score[2]
        if i == 5:
            p, pcls, score = p[:0], pcls[:0], score[:0]                     # an image without detections
        pb, tb = yxyx(p), yxyx(t)
        ref_map.cal_map(pb, pcls, score, np.zeros((138, 138, len(pb))), tb, tcls, np.zeros((138, 138, len(tb))), ap_data, iou_thresholds)
        out[f"p_bbox_{i}"], out[f"p_cls_{i}"], out[f"p_score_{i}"] = p, pcls, score
        out[f"t_bbox_{i}"], out[f"t_cls_{i}"] = t, tcls
    # per (threshold, class) AP of the box entries, then the reference's summary dict
    aps = np.full((len(iou_thresholds), ncls), -1.0)
    ngt = np.zeros((len(iou_thresholds), ncls), np.int64)
    npts = np.zeros((len(iou_thresholds), ncls), np.int64)
    for a, row in enumerate(ap_data["box"]):
        for c, obj in enumerate(row):
            ngt[a, c], npts[a, c] = obj.num_gt_positives, len(obj.data_points)
            if not obj.is_empty():
                aps[a, c] = obj.get_ap()
    maps = ref_map.calc_map(ap_data, iou_thresholds, class_names, print_result=False)
    out["box_ap"], out["box_ngt"], out["box_npts"] = aps, ngt, npts
    out["box_map_keys"] = np.array([str(k) for k in maps["box"].keys()])
    out["box_map_values"] = np.array(list(maps["box"].values()), np.float64)
    out["mask_map_values"] = np.array(list(maps["mask"].values()), np.float64)
    np.savez_compressed(os.path.join(HERE, "map_golden.npz"), **out)
    print("box mAP", dict(maps["box"]))
    print("mask mAP", dict(maps["mask"]))


if __name__ == "__main__":
    main()
