"""Mirror of detr_tf/inference.py:68-95 (get_model_inference) on device: softmax -> max score / argmax label ->
background filter -> box format, one launch of csrc/pipeline.cu:postprocess_kernel for the whole batch.
The drawing helper numpy_bbox_to_image (inference.py:11-65, cv2) is out of scope."""
import numpy as np
import torch

from . import ops

BBOX_FORMATS = {"xy_center": 0, "xyxy": 1, "yxyx": 2}


def _dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype).contiguous()


def batched_model_inference(m_outputs, background_class, bbox_format="xy_center", device="cuda"):
    """Extension (the reference handles image 0 only): every image of the batch.  Returns device tensors
    (boxes [B,Q,4] f32, labels [B,Q] i64, scores [B,Q] f32, query [B,Q] i32, count [B] i32); the first count[b] rows of image
    b are valid, in ascending query order.  No host sync."""
    if bbox_format not in BBOX_FORMATS:
        raise NotImplementedError()                      # inference.py:92-93
    lg = m_outputs["pred_logits"]
    device = lg.device if isinstance(lg, torch.Tensor) and lg.is_cuda else torch.device(device)
    logits = _dev(lg, torch.float32, device)
    boxes = _dev(m_outputs["pred_boxes"], torch.float32, device)
    B, Q, C = logits.shape
    out_boxes = torch.empty(B, Q, 4, dtype=torch.float32, device=device)
    out_labels = torch.empty(B, Q, dtype=torch.int64, device=device)
    out_scores = torch.empty(B, Q, dtype=torch.float32, device=device)
    out_query = torch.empty(B, Q, dtype=torch.int32, device=device)
    count = torch.empty(B, dtype=torch.int32, device=device)
    ops.postprocess(logits, C, boxes, B, Q, C, int(background_class), BBOX_FORMATS[bbox_format], out_boxes, out_labels,
                    out_scores, out_query, count)
    return out_boxes, out_labels, out_scores, out_query, count


def get_model_inference(m_outputs: dict, background_class, bbox_format="xy_center", device="cuda"):
    """inference.py:68-95: image 0 of the batch -> (predicted_bbox [k,4], predicted_labels [k] i64, predicted_scores [k]).
    The output length k depends on the data, so this wrapper reads one int back from the device (the reference's
    tf.where does the same)."""
    boxes, labels, scores, _, count = batched_model_inference(
        {"pred_logits": m_outputs["pred_logits"][:1], "pred_boxes": m_outputs["pred_boxes"][:1]}, background_class, bbox_format,
        device)
    k = int(count[0])
    return boxes[0, :k], labels[0, :k], scores[0, :k]
