"""Mirror of the two functions of detr_tf/data/processing.py that sit on the path into the train step:
normalized_images (:6-23) on device and pad_labels (:35-55), the T0 wire format.  The dataset loaders and the imgaug
augmentation pipeline (data/coco.py, voc.py, tfcsv.py, transformation.py) are out of scope."""
import numpy as np
import torch

from .. import ops


def normalisation_lut(normalized_method="torch_resnet"):
    """[3,256] float32 table: entry [c, v] = normalised value of byte v in OUTPUT channel c, computed with the reference's
    own float64 numpy arithmetic (processing.py:12-21) so the device lookup reproduces it bit for bit.
    Returns (lut, swap_rb): tf_resnet reverses the channel order (RGB -> BGR) before subtracting the caffe means."""
    v = np.arange(256, dtype=np.float64)[None, :]
    if normalized_method == "torch_resnet":
        channel_avg = np.array([0.485, 0.456, 0.406])[:, None]
        channel_std = np.array([0.229, 0.224, 0.225])[:, None]
        return ((v / 255.0 - channel_avg) / channel_std).astype(np.float32), False
    if normalized_method == "tf_resnet":
        mean = np.array([103.939, 116.779, 123.68])[:, None]
        return (v - mean).astype(np.float32), True
    raise Exception("Can't handler thid normalized method")          # processing.py:23 (sic)


_LUTS = {}


def device_lut(normalized_method, device):
    key = (normalized_method, str(device))
    if key not in _LUTS:
        lut, swap = normalisation_lut(normalized_method)
        _LUTS[key] = (torch.from_numpy(lut).to(device).contiguous(), swap)
    return _LUTS[key]


def normalized_images(image, config, device="cuda"):
    """processing.py:6-23 on device.  image: uint8 [H,W,3] or [B,H,W,3] (numpy or torch, host or device) -> float32 device
    tensor of the same shape.  (To skip the fp32 image altogether pass the uint8 batch straight to the model:
    `model(images_uint8, training=...)` fuses this normalisation into the stem's input layout.)"""
    if isinstance(image, np.ndarray):
        image = torch.from_numpy(image)
    if image.dtype != torch.uint8:
        raise TypeError("normalized_images: the device path takes uint8 pixels (0..255)")
    device = image.device if image.is_cuda else torch.device(device)
    img = image.to(device, non_blocking=True).contiguous()
    lut, swap = device_lut(config.normalized_method, device)
    out = torch.empty(img.shape, dtype=torch.float32, device=device)
    ops.normalize_u8(img, lut, swap, out, img.numel() // 3)
    return out


def pad_labels(images, t_bbox, t_class):
    """processing.py:35-55: ragged (n,4) boxes / (n,1) classes of one image -> the fixed 100-row wire format with a header
    row: t_bbox[0] = [n,0,0,0], t_class[0] = 0.  n <= 99 (the reference's tf.pad fails on a negative pad otherwise)."""
    tb = np.asarray(t_bbox.cpu() if isinstance(t_bbox, torch.Tensor) else t_bbox, dtype=np.float32).reshape(-1, 4)
    tc = np.asarray(t_class.cpu() if isinstance(t_class, torch.Tensor) else t_class, dtype=np.int64).reshape(-1, 1)
    n = tb.shape[0]
    if n > 99:
        raise ValueError(f"pad_labels: {n} boxes do not fit the 100-row wire format (max 99)")
    out_b = np.zeros((100, 4), np.float32)
    out_c = np.zeros((100, 1), np.int64)
    out_b[0, 0] = n
    out_b[1:1 + n] = tb
    out_c[1:1 + n] = tc
    return images, out_b, out_c
