"""Mirror of detr_tf/data/transformation.py on device: detr_aug_seq (:54-114) and detr_transform (:163-195).

Every augmenter the reference chains -- Fliplr(0.5), Sometimes(0.5, OneOf(Resize, CropToFixedSize, Affine(scale 0.5..1.5))) and
the final Resize to config.image_size -- is an axis-aligned affine map, so the whole sequence is ONE map per image: the host
draws it (a handful of random numbers) and pushes the boxes through it; the pixels are resampled once, bilinearly, by
`detrb_resize_affine_u8` (csrc/pipeline.cu) for the whole ragged batch in one launch, straight into the [B,H,W,3] uint8 batch
the model consumes (normalisation is fused into the stem's input layout).  Deviations from imgaug, stated: interpolation is
always bilinear (the reference draws one of nearest/linear/area/cubic per image, `ia.ALL`), one resampling instead of two, and
the random stream is numpy's, not imgaug's."""
import numpy as np
import torch

from .. import ops


def sample_geometry(src_hw, config, augmentation, rng=None):
    """transformation.py:54-95 for one source frame of size src_hw = (h, w) -> (fwd, zero_border):
    fwd = (fx, gx, fy, gy), output pixel coordinates x' = fx*x + gx, y' = fy*y + gy (continuous coordinates, pixel i covers
    [i, i+1)); zero_border: the map can look outside the source (Affine scale < 1: imgaug fills with cval = 0)."""
    h, w = float(src_hw[0]), float(src_hw[1])
    H, W = float(config.image_size[0]), float(config.image_size[1])
    fx, gx, fy, gy = 1.0, 0.0, 1.0, 0.0
    zero_border = False
    if augmentation:
        rng = rng or np.random.default_rng()
        if rng.random() < 0.5:                                  # iaa.Fliplr(0.5)
            fx, gx = -fx, w - gx
        if rng.random() < 0.5:                                  # sometimes(OneOf([...]))
            which = int(rng.integers(0, 3))
            if which == 0:                                      # Resize to the target size
                fx, gx, fy, gy = fx * W / w, gx * W / w, fy * H / h, gy * H / h
                w, h = W, H
            elif which == 1:                                    # CropToFixedSize(W, H): crops only the axes that are larger
                ux, uy = rng.random(), rng.random()
                if w > W:
                    gx -= float(int(ux * (w - W)))
                    w = W
                if h > H:
                    gy -= float(int(uy * (h - H)))
                    h = H
            else:                                               # Affine(scale x,y in (0.5, 1.5)) about the image centre
                sx, sy = rng.uniform(0.5, 1.5), rng.uniform(0.5, 1.5)
                fx, gx = fx * sx, sx * (gx - w / 2) + w / 2
                fy, gy = fy * sy, sy * (gy - h / 2) + h / 2
                zero_border = True
    fx, gx, fy, gy = fx * W / w, gx * W / w, fy * H / h, gy * H / h     # final Resize (both branches, :79 / :88)
    return (fx, gx, fy, gy), zero_border


def transform_boxes(bbox, t_class, fwd, out_hw, src_hw):
    """transformation.py:11-34, :117-142, :178-186 vectorised: normalised (xc,yc,w,h) -> pixel corners -> the map ->
    remove_out_of_image_fraction(0.7) -> clip_out_of_image -> normalised (xc,yc,w,h) of the output frame."""
    b = np.asarray(bbox, np.float64).reshape(-1, 4)
    c = np.asarray(t_class).reshape(-1)
    sh, sw = src_hw
    H, W = out_hw
    fx, gx, fy, gy = fwd
    xa, xb = fx * (b[:, 0] - b[:, 2] / 2) * sw + gx, fx * (b[:, 0] + b[:, 2] / 2) * sw + gx
    ya, yb = fy * (b[:, 1] - b[:, 3] / 2) * sh + gy, fy * (b[:, 1] + b[:, 3] / 2) * sh + gy
    x1, x2, y1, y2 = np.minimum(xa, xb), np.maximum(xa, xb), np.minimum(ya, yb), np.maximum(ya, yb)
    area = (x2 - x1) * (y2 - y1)
    ix1, ix2, iy1, iy2 = np.clip(x1, 0, W), np.clip(x2, 0, W), np.clip(y1, 0, H), np.clip(y2, 0, H)
    inside = (ix2 - ix1) * (iy2 - iy1)
    point_in = (x1 >= 0) & (x1 < W) & (y1 >= 0) & (y1 < H)
    frac_out = np.where(area > 0, 1.0 - inside / np.where(area > 0, area, 1.0), np.where(point_in, 0.0, 1.0))
    keep = frac_out < 0.7
    w, h = (ix2 - ix1)[keep], (iy2 - iy1)[keep]
    out = np.stack([(ix1[keep] + w / 2) / W, (iy1[keep] + h / 2) / H, w / W, h / H], -1)
    return out.reshape(-1, 4), c[keep]


def inverse_map(fwd):
    """(fx,gx,fy,gy) -> the kernel's (ax,bx,ay,by): source coordinate of an output coordinate"""
    fx, gx, fy, gy = fwd
    return (1.0 / fx, -gx / fx, 1.0 / fy, -gy / fy)


def resample_batch(frames, fwds, zero_borders, out_hw, device="cuda"):
    """frames: list of uint8 [h,w,3] arrays (numpy or torch, ragged) -> uint8 DEVICE batch [B,H,W,3], one kernel launch."""
    B = len(frames)
    H, W = int(out_hw[0]), int(out_hw[1])
    flat, off, hw = [], [], []
    n = 0
    for f in frames:
        t = torch.from_numpy(np.ascontiguousarray(f)) if isinstance(f, np.ndarray) else f
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise TypeError("resample_batch: frames are uint8 [h,w,3]")
        off.append(n)
        hw += [t.shape[0], t.shape[1]]
        flat.append(t.reshape(-1))
        n += (t.numel() + 15) // 16 * 16
    device = torch.device(device)
    src = torch.zeros(n, dtype=torch.uint8, device=device)
    for t, o in zip(flat, off):
        src[o:o + t.numel()].copy_(t, non_blocking=True)
    inv = torch.tensor([inverse_map(f) for f in fwds], dtype=torch.float32).to(device)
    out = torch.empty(B, H, W, 3, dtype=torch.uint8, device=device)
    ops.resize_affine_u8(src, torch.tensor(off, dtype=torch.int64).to(device), torch.tensor(hw, dtype=torch.int32).to(device), inv,
                         torch.tensor([1 if z else 0 for z in zero_borders], dtype=torch.uint8).to(device), out, B, H, W)
    return out


def detr_transform_batch(images, bboxes, classes, config, augmentation, rng=None, device="cuda"):
    """Batched detr_transform: (uint8 device batch [B,H,W,3], list of (n_i,4) boxes, list of (n_i,) classes)."""
    out_hw = (int(config.image_size[0]), int(config.image_size[1]))
    fwds, zbs, ob, oc = [], [], [], []
    for img, bb, cc in zip(images, bboxes, classes):
        fwd, zb = sample_geometry(img.shape[:2], config, augmentation, rng)
        nb, nc = transform_boxes(bb, cc, fwd, out_hw, img.shape[:2])
        fwds.append(fwd), zbs.append(zb), ob.append(nb), oc.append(nc)
    return resample_batch(images, fwds, zbs, out_hw, device), ob, oc


def detr_transform(image, bbox, t_class, config, augmentation, rng=None, device="cuda"):
    """transformation.py:163-195, same signature and return order: (image float32 [H,W,3], bbox (n,4), t_class (n,)).  The
    image stays on the device (float32 like the reference's `astype(np.float32)`, values 0..255)."""
    imgs, ob, oc = detr_transform_batch([image], [bbox], [t_class], config, augmentation, rng, device)
    return imgs[0].to(torch.float32), ob[0], oc[0]
