"""Pins the conventions of the TMA im2col tensor map (cuTensorMapEncodeIm2col + cp.async.bulk.tensor.4d.im2col) that the
tcgen05 convolution kernel relies on: bounding-box corners, traversal stride, start coordinate, filter offsets, zero fill,
pixel order across rows / images, and the 128-byte swizzle of the landed tile."""
import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


_PROBE = None


def _probe_lib():
    """tests/_harness_build/libdetrb_probe.so (tests/native/tma_probe.cu, built by __graft_entry__.build()): test infrastructure,
    linked against the product library"""
    global _PROBE
    if _PROBE is None:
        import os
        from detr_tensorflow_b200 import _lib
        _lib.lib()                                                   # libdetrb.so first (the probe library resolves its symbols from it)
        here = os.path.dirname(os.path.abspath(__file__))
        so = os.path.join(here, "_harness_build", "libdetrb_probe.so")
        if not os.path.exists(so):
            import sys
            sys.path.insert(0, os.path.dirname(here))
            import __graft_entry__ as g
            g.build_test_probe(_lib._SO)
        _PROBE = ctypes.CDLL(so, mode=ctypes.RTLD_GLOBAL)
    return _PROBE


def probe(x, lower, upper, stride, pixels, swz, c0, w, h, n, off_w, off_h):
    from detr_tensorflow_b200 import _lib
    B, H, W, C = x.shape
    out = torch.zeros(pixels * 128 + 1, dtype=torch.uint8, device="cuda")
    _lib.check(_probe_lib().detrb_tma_im2col_probe(ctypes.c_void_p(x.data_ptr()), B, H, W, C, lower[0], lower[1], upper[0], upper[1],
                                                 stride, pixels, swz, c0, w, h, n, off_w, off_h, ctypes.c_void_p(out.data_ptr()),
                                                 ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    raw = out.cpu()
    done = int(raw[-1])
    tile = raw[:-1].view(torch.bfloat16).view(pixels, 64).clone()
    if swz:      # undo the 128B swizzle: 16-byte chunk c of row r sits at chunk c ^ (r % 8)
        t = tile.view(pixels, 8, 8)
        un = torch.empty_like(t)
        for r in range(pixels):
            for c in range(8):
                un[r, c] = t[r, c ^ (r % 8)]
        tile = un.view(pixels, 64)
    return done, tile


def expected(x, lower, stride, pixels, c0, ox0, oy0, n0, OW, OH, off_w, off_h):
    B, H, W, C = x.shape
    out = torch.zeros(pixels, 64, dtype=torch.bfloat16)
    ox, oy, n = ox0, oy0, n0
    for p in range(pixels):
        if n < B:
            ix, iy = lower[0] + ox * stride + off_w, lower[1] + oy * stride + off_h
            if 0 <= ix < W and 0 <= iy < H:
                out[p] = x[n, iy, ix, c0:c0 + 64].cpu()
        ox += 1
        if ox == OW:
            ox, oy = 0, oy + 1
            if oy == OH:
                oy, n = 0, n + 1
    return out


def locate(x, row):
    """which pixel of x does this 64-vector come from (c0 = 0)? -> (n, y, x) or 'zero' / '?'"""
    if float(row.float().abs().max()) == 0:
        return "zero"
    B, H, W, C = x.shape
    flat = x[..., :64].reshape(-1, 64).cpu()
    hit = (flat == row).all(dim=1).nonzero()
    if len(hit) == 0:
        return "?"
    i = int(hit[0])
    return (i // (H * W), (i // W) % H, i % W)


CASES = [  # name, H, W, k, stride, pad, start (ox, oy, n), tap (kw, kh)
    ("3x3s1 interior tap", 9, 13, 3, 1, 1, (5, 2, 0), (1, 1)),
    ("3x3s1 corner tap wrap", 9, 13, 3, 1, 1, (9, 7, 0), (0, 0)),
    ("3x3s1 last tap", 9, 13, 3, 1, 1, (0, 0, 1), (2, 2)),
    ("3x3s2", 9, 13, 3, 2, 1, (3, 1, 0), (2, 0)),
    ("1x1s2", 9, 13, 1, 2, 0, (2, 3, 0), (0, 0)),
    ("1x1s1", 9, 13, 1, 1, 0, (4, 8, 0), (0, 0)),
]


@pytest.mark.parametrize("swz", [0, 1])
@pytest.mark.parametrize("name,H,W,k,stride,pad,start,tap", CASES)
def test_im2col_conventions(name, H, W, k, stride, pad, start, tap, swz):
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    B, C, pixels = 2, 128, 32
    x = (torch.arange(B * H * W * C, dtype=torch.float32).reshape(B, H, W, C) % 251 + torch.arange(B * H * W).reshape(B, H, W, 1) * 0.5)
    x = x.to(torch.bfloat16).cuda()
    OW, OH = (W + 2 * pad - k) // stride + 1, (H + 2 * pad - k) // stride + 1
    lower = (-pad, -pad)
    upper = (pad - (k - 1), pad - (k - 1))
    ox0, oy0, n0 = start
    done, tile = probe(x, lower, upper, stride, pixels, swz, 0, lower[0] + ox0 * stride, lower[1] + oy0 * stride, n0, tap[0], tap[1])
    exp = expected(x, lower, stride, pixels, 0, ox0, oy0, n0, OW, OH, tap[0], tap[1])
    ok = torch.equal(tile, exp)
    if not ok:
        got = [locate(x, tile[p]) for p in range(12)]
        want = [locate(x, exp[p]) for p in range(12)]
        pytest.fail(f"{name} swz={swz} done={done}: rows map to {got}, expected {want}")
    assert done == 1
