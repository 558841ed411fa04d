"""Golden vectors for the rows either side of the train step, by EXECUTING THE REFERENCE'S OWN CODE:

    python tests/golden/make_golden_pipeline.py      # needs /root/reference (build container only)

detr_tf/inference.py:get_model_inference (:68-95), detr_tf/data/processing.py:normalized_images (:6-23) and
pad_labels (:35-55) are imported unmodified from /root/reference on the numpy `tensorflow` shim (cv2 stubbed: only
the drawing helper uses it).  Output -> tests/golden/pipeline_golden.npz (committed).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_numpy_shim  # noqa: E402

tf_numpy_shim.install()
sys.modules.setdefault("cv2", types.ModuleType("cv2"))
sys.path.insert(0, "/root/reference")
from detr_tf import inference as ref_inf  # noqa: E402
import importlib.util  # noqa: E402

# detr_tf/data/__init__.py imports the dataset loaders (pycocotools, imgaug: missing); load processing.py itself, unmodified
_spec = importlib.util.spec_from_file_location("ref_processing", "/root/reference/detr_tf/data/processing.py")
ref_proc = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(ref_proc)


class Cfg:
    normalized_method = "torch_resnet"


def main():
    out = {}
    rs = np.random.RandomState(4321)
    # ---- N2: get_model_inference.  Peaked logits so that a realistic share of queries is non-background; ties on purpose.
    B, Q, C = 3, 100, 92
    logits = rs.randn(B, Q, C).astype(np.float32)
    logits[:, :, 91] += 2.5                                   # most queries -> background (91)
    hot = rs.rand(B, Q) < 0.3
    cls = rs.randint(0, 91, (B, Q))
    for b in range(B):
        for q in range(Q):
            if hot[b, q]:
                logits[b, q, cls[b, q]] += 6.0
    logits[0, 7, :] = 0.0                                     # full tie: argmax -> class 0 (first index)
    logits[0, 8, 5] = logits[0, 8, 91] = 9.0                  # two-way tie incl. background: first index (5) wins
    boxes = np.concatenate([rs.uniform(0.05, 0.95, (B, Q, 2)), rs.uniform(0.02, 0.9, (B, Q, 2))], -1).astype(np.float32)
    out["i_logits"], out["i_boxes"] = logits, boxes
    for bg in (91, 0):
        for fmt in ("xy_center", "xyxy", "yxyx"):
            for b in range(B):
                m = {"pred_logits": logits[b:b + 1], "pred_boxes": boxes[b:b + 1]}
                pb, pl, ps = ref_inf.get_model_inference(m, bg, bbox_format=fmt)
                out[f"i_{bg}_{fmt}_{b}_bbox"] = np.asarray(pb, np.float32)
                out[f"i_{bg}_{fmt}_{b}_labels"] = np.asarray(pl, np.int64)
                out[f"i_{bg}_{fmt}_{b}_scores"] = np.asarray(ps, np.float32)
    # ---- N1: normalized_images (uint8 pixels -> float32) for both methods; every byte value in every channel
    img = rs.randint(0, 256, (2, 37, 53, 3)).astype(np.uint8)
    img[0, 0, :, :] = 0
    img[0, 1, :, :] = 255
    ramp = np.arange(256, dtype=np.uint8)
    img[1, 2, :, :] = ramp[:53, None]
    out["n_img"] = img
    for method in ("torch_resnet", "tf_resnet"):
        cfg = Cfg()
        cfg.normalized_method = method
        out[f"n_{method}"] = np.stack([ref_proc.normalized_images(img[b], cfg) for b in range(2)])
        lut = ref_proc.normalized_images(np.repeat(ramp[:, None], 3, 1).reshape(256, 1, 3), cfg)
        out[f"n_{method}_lut"] = lut.reshape(256, 3)
    # ---- T0: pad_labels
    for k, n in enumerate((0, 1, 20, 99)):
        tb = rs.uniform(0.0, 1.0, (n, 4)).astype(np.float32)
        tc = rs.randint(0, 91, (n, 1)).astype(np.int64)
        _, pb, pc = ref_proc.pad_labels(None, tb, tc)
        out[f"p_{k}_in_bbox"], out[f"p_{k}_in_class"] = tb, tc
        out[f"p_{k}_bbox"], out[f"p_{k}_class"] = np.asarray(pb, np.float32), np.asarray(pc, np.int64)
    path = os.path.join(HERE, "pipeline_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")
    for b in range(B):
        print("image", b, "kept", len(out[f"i_91_xyxy_{b}_labels"]), "of", Q)


if __name__ == "__main__":
    main()
