"""CPU restatement of the input-geometry step -- TEST INFRASTRUCTURE (the checker), never the product path.

Follows /root/reference/detr_tf/data/transformation.py:54-114 (detr_aug_seq: Fliplr, Sometimes(OneOf(Resize, CropToFixedSize,
Affine scale)), final Resize to config.image_size) and :163-195 (detr_transform: boxes through the same geometry,
remove_out_of_image_fraction(0.7), clip_out_of_image, back to normalised xc,yc,w,h).  The arithmetic lives in imgaug 0.4.0
(requirements.txt: `imgaug`, unpinned; NOT installed here and not vendored under /root/reference): every augmenter of that
sequence is an axis-aligned affine map, restated here as one composed map per image with cv2.INTER_LINEAR's published
pixel-centre convention.  PARITY UNPINNED against imgaug itself (its interpolation is drawn from ia.ALL = nearest / linear /
area / cubic at random: the device path always resamples bilinearly -- a documented deviation); the plain bilinear resize is
pinned against torch.nn.functional.interpolate(align_corners=False), the same convention, in tests/test_oracle_cpu.py.
"""
import numpy as np


def resize_affine_u8(frames, inv, zero_border, H, W):
    """frames: list of uint8 [h,w,3]; inv [B,4] float32 = (ax, bx, ay, by); zero_border [B] -> uint8 [B,H,W,3].
    Operation order and rounding identical to csrc/pipeline.cu:resize_affine_u8_kernel (all float32, round-to-nearest)."""
    f32 = np.float32
    B = len(frames)
    out = np.zeros((B, H, W, 3), np.uint8)
    xc = np.arange(W, dtype=f32) + f32(0.5)
    yc = np.arange(H, dtype=f32) + f32(0.5)
    for b in range(B):
        img = frames[b]
        sh, sw = img.shape[:2]
        ax, bx, ay, by = (f32(v) for v in inv[b])
        xs = ((ax * xc).astype(f32) + bx).astype(f32) - f32(0.5)
        ys = ((ay * yc).astype(f32) + by).astype(f32) - f32(0.5)
        x0f, y0f = np.floor(xs), np.floor(ys)
        fx, fy = (xs - x0f).astype(f32), (ys - y0f).astype(f32)
        x0, y0 = x0f.astype(np.int64), y0f.astype(np.int64)
        acc = np.zeros((H, W, 3), f32)
        for dy in (0, 1):
            wy = fy if dy else (f32(1) - fy).astype(f32)
            yy = y0 + dy
            for dx in (0, 1):
                wx = fx if dx else (f32(1) - fx).astype(f32)
                xx = x0 + dx
                w = (wx[None, :] * wy[:, None]).astype(f32)
                inside = ((xx >= 0) & (xx < sw))[None, :] & ((yy >= 0) & (yy < sh))[:, None]
                px = img[np.clip(yy, 0, sh - 1)[:, None], np.clip(xx, 0, sw - 1)[None, :]].astype(f32)
                term = (w[..., None] * px).astype(f32)
                if zero_border[b]:
                    term = np.where(inside[..., None], term, f32(0))
                acc = (acc + term).astype(f32)
        out[b] = np.clip(np.rint(acc), 0, 255).astype(np.uint8)
    return out


def transform_boxes(bbox, t_class, fwd, out_hw, src_hw):
    """transformation.py:163-195 on the boxes of one image: normalised (xc,yc,w,h) of the source frame -> pixel corners
    (:11-34) -> forward map x' = fx*x + gx, y' = fy*y + gy (fwd = (fx, gx, fy, gy); a flip has fx < 0, so corners are
    re-sorted as imgaug's BoundingBox does) -> drop boxes whose out-of-image area fraction is >= 0.7 -> clip to the output
    frame -> normalised (xc,yc,w,h) (:117-142).  float64 like the reference's python floats."""
    sh, sw = src_hw
    H, W = out_hw
    fx, gx, fy, gy = (float(v) for v in fwd)
    keep_b, keep_c = [], []
    for bb, c in zip(np.asarray(bbox, np.float64).reshape(-1, 4), np.asarray(t_class).reshape(-1)):
        xcs, ycs, ws, hs = bb[0] * sw, bb[1] * sh, bb[2] * sw, bb[3] * sh
        x1, x2 = fx * (xcs - ws / 2) + gx, fx * (xcs + ws / 2) + gx
        y1, y2 = fy * (ycs - hs / 2) + gy, fy * (ycs + hs / 2) + gy
        x1, x2 = min(x1, x2), max(x1, x2)
        y1, y2 = min(y1, y2), max(y1, y2)
        area = (x2 - x1) * (y2 - y1)
        ix1, ix2, iy1, iy2 = min(max(x1, 0.0), W), min(max(x2, 0.0), W), min(max(y1, 0.0), H), min(max(y2, 0.0), H)
        inside = max(ix2 - ix1, 0.0) * max(iy2 - iy1, 0.0)
        frac_out = 1.0 - inside / area if area > 0 else (0.0 if (0 <= x1 < W and 0 <= y1 < H) else 1.0)
        if frac_out >= 0.7:
            continue
        w, h = ix2 - ix1, iy2 - iy1
        keep_b.append([(ix1 + w / 2) / W, (iy1 + h / 2) / H, w / W, h / H])
        keep_c.append(c)
    return np.asarray(keep_b, np.float64).reshape(-1, 4), np.asarray(keep_c)
