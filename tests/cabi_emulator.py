"""A pure-PyTorch (CPU) stand-in for libdetrb.so's C ABI -- TEST INFRASTRUCTURE ONLY.

Purpose: exercise the Python orchestration of the engine (buffer plumbing, layer order, hand-written
backward chain, parameter layouts, optimizer glue) against the oracle on a machine without a GPU.  Each
function re-implements the *documented semantics* of the matching entry point of include/detrb.h on raw
host pointers (bf16 storage, fp32 accumulate).  It says nothing about the CUDA kernels themselves -- those
are checked against the oracle on a B200 by the `-m gpu` tests.  The product never imports this file.
"""
import ctypes
import os
import subprocess

import numpy as np
import torch
import torch.nn.functional as F

from detr_tensorflow_b200 import _lib

F32 = torch.float32


class _Act:
    """storage dtype standing in for the device's bf16 tensors; tests may switch it to float32 (together with
    engine.BF16) to check the orchestration without rounding noise"""
    dtype = torch.bfloat16


def set_act_dtype(dt):
    _Act.dtype = dt
    import detr_tensorflow_b200.engine as E
    E.BF16 = dt

_ITEM = {torch.bfloat16: 2, F32: 4, torch.int64: 8, torch.int32: 4, torch.uint8: 1, torch.float64: 8}


def _addr(p):
    if p is None:
        return 0
    if isinstance(p, ctypes.c_void_p):
        return p.value or 0
    if isinstance(p, int):
        return p
    return getattr(p, "value", 0) or 0


def T(p, dtype, count):
    """1-D tensor view of `count` elements of `dtype` at raw host address p (None if NULL)."""
    a = _addr(p)
    if a == 0:
        return None
    buf = (ctypes.c_uint8 * (count * _ITEM[dtype])).from_address(a)
    return torch.frombuffer(buf, dtype=dtype, count=count)


T_ = T        # (alias: some emulator methods have a parameter called T)


def M2(p, dtype, rows, cols, ld):
    """[rows, cols] strided view with row stride ld"""
    t = T(p, dtype, (rows - 1) * ld + cols)
    return None if t is None else torch.as_strided(t, (rows, cols), (ld, 1))


def _v(x):
    return x.value if hasattr(x, "value") else x


# ---- parity precision (include/detrb.h): an activation element is a PAIR of storage values, hi at p and lo = x - hi at
#      p + split elements; split == 0: plain storage
def _lo(p, split):
    return _addr(p) + int(_v(split)) * _ITEM[_Act.dtype]


def RD(p, rows, cols, ld, split=0):
    """fp32 [rows, cols] value of a (possibly paired) activation matrix"""
    v = M2(p, _Act.dtype, rows, cols, ld).to(F32)
    if _v(split):
        v = v + M2(_lo(p, split), _Act.dtype, rows, cols, ld).to(F32)
    return v


def WR(p, rows, cols, ld, val, split=0, index=None):
    """store fp32 val as the (pair of) storage values; index: row subset"""
    hi = val.to(_Act.dtype)
    dst = M2(p, _Act.dtype, rows, cols, ld)
    if index is None:
        dst[:] = hi
    else:
        dst[index] = hi
    if _v(split):
        lo = (val - hi.to(F32)).to(_Act.dtype)
        dst = M2(_lo(p, split), _Act.dtype, rows, cols, ld)
        if index is None:
            dst[:] = lo
        else:
            dst[index] = lo


def RND(val, split=0):
    """the value a later RD of a WR(val) returns"""
    hi = val.to(_Act.dtype).to(F32)
    if _v(split):
        return hi + (val - hi).to(_Act.dtype).to(F32)
    return hi


# ---------------------------------------------------------------- dropout hash (mirrors csrc/common.cuh)
def _lowbias32(x):
    x = x.astype(np.uint64)
    x ^= x >> 16
    x = (x * 0x7feb352d) & 0xffffffff
    x ^= x >> 15
    x = (x * 0x846ca68b) & 0xffffffff
    x ^= x >> 16
    return x


def keep_mask(rows, cols, drop_p, seed, site, seed_ptr=None):
    """bool [len(rows), len(cols)] keep mask for element (row, col)"""
    seed = int(seed)
    if _addr(seed_ptr):
        seed ^= int(T(seed_ptr, torch.int64, 1)[0]) & 0xffffffffffffffff
    r = np.asarray(rows, dtype=np.uint64)[:, None]
    c = np.asarray(cols, dtype=np.uint64)[None, :]
    rowhash = _lowbias32((r ^ (seed & 0xffffffff) ^ ((site * 0x9E3779B9) & 0xffffffff)) & 0xffffffff) ^ (seed >> 32)
    h = ((rowhash ^ (((c >> 1) * 0x85EBCA77) & 0xffffffff)) & 0xffffffff).astype(np.uint64)   # dropout_bits_rh (common.cuh)
    h ^= h >> 16
    h = (h * 0x7feb352d) & 0xffffffff
    h ^= h >> 15
    bits = np.where((c & 1) == 1, h >> 16, h & 0xffff)
    thresh = int(drop_p * 65536.0 + 0.5)
    return torch.from_numpy(bits >= thresh)


def attn_keep_mask(rows, keys, drop_p, seed, site, seed_ptr=None):
    """bool [len(rows), len(keys)]: attention-probability dropout (attn_drop_word / attn_drop_keepbits, csrc/common.cuh)"""
    seed = int(seed)
    if _addr(seed_ptr):
        seed ^= int(T(seed_ptr, torch.int64, 1)[0]) & 0xffffffffffffffff
    r = np.asarray(rows, dtype=np.uint64)[:, None]
    k = np.asarray(keys, dtype=np.uint64)[None, :]
    rowhash = _lowbias32((r ^ (seed & 0xffffffff) ^ ((site * 0x9E3779B9) & 0xffffffff)) & 0xffffffff) ^ (seed >> 32)
    rowhash = (rowhash * 0x21F0AAAD) & 0xffffffff
    pair = ((k >> 4) << 3) | (k & 7)
    x = (rowhash + pair * ((0x9E3779B1 * 0x21F0AAAD) & 0xffffffff)) & 0xffffffff
    x ^= x >> 15
    x = (x * 0x735A2D97) & 0xffffffff
    field = np.where(((k >> 3) & 1) == 1, (x >> 16) & 0x7fff, x & 0x7fff)
    return torch.from_numpy(field >= int(np.float32(drop_p) * np.float32(32768.0) + np.float32(0.5)))


# ---------------------------------------------------------------- gather
def _gather(p_A, lda, M, K, batch, IH, IW, Cin, OH, OW, KH, KW, stride, pad, mode, stem_real_kw=None):
    A = M2(p_A, _Act.dtype, batch * IH * IW, Cin, lda).to(F32)
    m = torch.arange(M)
    b = m // (OH * OW)
    rem = m % (OH * OW)
    oy, ox = rem // OW, rem % OW
    cols = []
    for kh in range(KH):
        for kw in range(KW):
            if mode == 0:
                iy, ix = oy * stride - pad + kh, ox * stride - pad + kw
                ok = torch.ones(M, dtype=torch.bool)
            else:
                ty, tx = oy + pad - kh, ox + pad - kw
                ok = (ty >= 0) & (tx >= 0) & (ty % stride == 0) & (tx % stride == 0)
                iy, ix = torch.div(ty, stride, rounding_mode="floor"), torch.div(tx, stride, rounding_mode="floor")
            ok = ok & (iy >= 0) & (iy < IH) & (ix >= 0) & (ix < IW)
            if stem_real_kw is not None and kw >= stem_real_kw:
                ok = ok & False
            idx = (b * IH + iy.clamp(0, IH - 1)) * IW + ix.clamp(0, IW - 1)
            cols.append(A[idx] * ok[:, None].to(F32))
    out = torch.cat(cols, dim=1)
    assert out.shape[1] == K, (out.shape, K)
    return out


def _sliding(p_A, lda, M, K, kb_rows):
    """sliding-window A operand (detrb_igemm_t.a_kb_rows): k-block j of row m = 64 elements at A + (m + j*kb_rows)*lda"""
    nk = K // 64
    flat = T(p_A, _Act.dtype, (M + (nk - 1) * kb_rows) * lda + 64).to(F32)
    m = torch.arange(M)
    cols = []
    for j in range(nk):
        idx = ((m + j * kb_rows) * lda)[:, None] + torch.arange(64)[None, :]
        cols.append(flat[idx])
    return torch.cat(cols, dim=1)


class FakeLib:
    def __init__(self):
        self.err = b""
        self._harness = None

    # -- plumbing
    def detrb_version(self):
        return _lib.ABI_VERSION

    def detrb_last_error(self):
        return self.err

    def detrb_check_device(self):
        return 0

    def harness(self):
        if self._harness is None:
            root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
            so = os.path.join(root, "tests", "_harness_build", "libharness.so")
            os.makedirs(os.path.dirname(so), exist_ok=True)
            subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                                   os.path.join(root, "tests", "host_harness.c"), "-lm"])
            self._harness = ctypes.CDLL(so)
        return self._harness

    # -- GEMM / conv
    def detrb_igemm(self, pref, stream):
        p = pref._obj
        stem = p.Cin == 4
        def gathered(addr):
            if p.a_kb_rows:
                return _sliding(addr, p.lda, p.M, p.K, p.a_kb_rows)
            return _gather(addr, p.lda, p.M, p.K, p.batch, p.IH, p.IW, p.Cin, p.OH, p.OW, p.KH, p.KW, p.stride, p.pad, p.mode)
        Ag = gathered(p.A)
        W = M2(p.W, _Act.dtype, p.N, p.K, p.ldw).to(F32)
        v = Ag @ W.t()
        sp = int(p.split)
        if sp:                           # three products: hi*hi + lo*hi + hi*lo
            assert p.wsplit
            Al = gathered(_lo(p.A, sp))
            Wl = M2(_lo(p.W, p.wsplit), _Act.dtype, p.N, p.K, p.ldw).to(F32)
            v = v + Al @ W.t() + Ag @ Wl.t()
        out_stride = max(p.out_stride, 1)
        m = torch.arange(p.M)
        if out_stride > 1:
            b = m // (p.OH * p.OW)
            rem = m % (p.OH * p.OW)
            orow = (b * p.SH + (rem // p.OW) * out_stride) * p.SW + (rem % p.OW) * out_stride
            nrows = p.batch * p.SH * p.SW
        else:
            orow, nrows = m, p.M
        if _addr(p.bias):
            v = v + T(p.bias, F32, p.N)
        res = RD(p.residual, nrows, p.N, p.ldr, sp)[orow] if _addr(p.residual) else None
        if res is not None and not p.drop_p > 0:
            v = v + res
        if p.relu:
            v = F.relu(v)
        if _addr(p.mask):
            mk = M2(p.mask, _Act.dtype, nrows, p.N, p.ldm)[orow].to(F32)
            v = torch.where(mk > 0, v * p.mask_scale, torch.zeros_like(v))
        if _addr(p.mask_bits):
            assert not _addr(p.mask) and p.N % 64 == 0
            by = M2(p.mask_bits, torch.uint8, nrows, p.N // 8, p.ldmb)[orow]
            bit = (by[:, :, None] >> torch.arange(8, dtype=torch.uint8)[None, None, :]) & 1
            v = torch.where(bit.reshape(p.M, p.N) > 0, v * p.mask_scale, torch.zeros_like(v))
        if p.sigmoid:
            v = torch.sigmoid(v)
        if p.drop_p > 0:
            keep = keep_mask(range(p.M), range(p.N), p.drop_p, p.seed, p.site, p.seed_ptr)
            v = torch.where(keep, v / (1 - p.drop_p), torch.zeros_like(v))
            if res is not None:
                v = v + res
        if _addr(p.C):
            if p.accumulate:
                v = v + RD(p.C, nrows, p.N, p.ldc, sp)[orow]
            WR(p.C, nrows, p.N, p.ldc, v, sp, index=orow)
        if _addr(p.Cf):
            M2(p.Cf, F32, nrows, p.N, p.ldcf)[orow] = v
        if _addr(p.out_bits):
            assert p.N % 64 == 0
            w = (1 << torch.arange(8, dtype=torch.int32))[None, None, :]
            by = ((v > 0).reshape(p.M, p.N // 8, 8).to(torch.int32) * w).sum(-1).to(torch.uint8)
            M2(p.out_bits, torch.uint8, nrows, p.N // 8, p.ldob)[orow] = by
        return 0

    def detrb_wgrad(self, pref, stream):
        p = pref._obj
        stem = p.Cin == 4
        def gathered(addr):
            if p.a_kb_rows:
                return _sliding(addr, p.lda, p.M, p.K, p.a_kb_rows)
            return _gather(addr, p.lda, p.M, p.K, p.batch, p.IH, p.IW, p.Cin, p.OH, p.OW, p.KH, p.KW, p.stride, p.pad, 0,
                           stem_real_kw=7 if stem else None)
        Ag = gathered(p.A)
        dY = M2(p.dY, _Act.dtype, p.M, p.N, p.ldy).to(F32)
        g = dY.t() @ Ag
        sp = int(p.split)
        if sp:
            dYl = M2(_lo(p.dY, sp), _Act.dtype, p.M, p.N, p.ldy).to(F32)
            g = g + dYl.t() @ Ag + dY.t() @ gathered(_lo(p.A, sp))
            dY = dY + dYl
        if p.k_mask or (p.Cin == 16 and p.KH == 4 and p.KW == 4 and p.pad == 2):
            # space-to-depth stem: columns of taps that do not exist in the 7x7x3 kernel get no gradient
            k = torch.arange(p.K)
            ch, tb, ta = k % 16, (k // 16) % 4, k // 64
            ry, rx = ch // 6, (ch // 3) % 2
            kh, kw = 2 * ta + ry - 1, 2 * tb + rx - 1
            g = g * ((ch < 12) & (kh >= 0) & (kh <= 6) & (kw >= 0) & (kw <= 6)).to(F32)[None, :]
        sc = T(p.rowscale, F32, p.N) if _addr(p.rowscale) else None
        if sc is not None:
            g = g * sc[:, None]
        M2(p.dW, F32, p.N, p.K, p.ldw).add_(g)
        if _addr(p.dbias):
            s = dY.sum(0)
            T(p.dbias, F32, p.N).add_(s * sc if sc is not None else s)
        return 0

    # -- attention
    @staticmethod
    def _attn_core(Q, K, V, p, B, H, Lq, Lk):
        """Q [B,Lq,H*32] etc float32 -> O, lse; differentiable"""
        q = Q.view(B, Lq, H, 32).transpose(1, 2)
        k = K.view(B, Lk, H, 32).transpose(1, 2)
        v = V.view(B, Lk, H, 32).transpose(1, 2)
        s = (q @ k.transpose(-1, -2)) * p.scale
        lse = torch.logsumexp(s, -1)
        w = torch.softmax(s, -1)
        if p.drop_p > 0:
            keep = attn_keep_mask(range(B * H * Lq), range(Lk), p.drop_p, p.seed, p.site, p.seed_ptr).view(B, H, Lq, Lk)
            w = torch.where(keep, w / (1 - p.drop_p), torch.zeros_like(w))
        if not p.split:
            w = w.to(_Act.dtype).to(F32)           # the tensor-core kernels feed bf16 probabilities to the P.V product
        o = (w @ v).transpose(1, 2).reshape(B, Lq, H * 32)
        return o, lse

    def detrb_attn_fwd(self, pref, stream):
        p = pref._obj
        B, H, Lq, Lk = p.B, p.H, p.Lq, p.Lk
        sp = int(p.split)
        Q = RD(p.Q, B * Lq, H * 32, p.ldq, sp).view(B, Lq, -1)
        K = RD(p.K, B * Lk, H * 32, p.ldk, sp).view(B, Lk, -1)
        V = RD(p.V, B * Lk, H * 32, p.ldv, sp).view(B, Lk, -1)
        o, lse = self._attn_core(Q, K, V, p, B, H, Lq, Lk)
        WR(p.O, B * Lq, H * 32, p.ldo, o.reshape(B * Lq, -1), sp)
        if _addr(p.lse):
            T(p.lse, F32, B * H * Lq)[:] = lse.reshape(-1)
        return 0

    def detrb_attn_bwd(self, pref, stream):
        p = pref._obj
        B, H, Lq, Lk = p.B, p.H, p.Lq, p.Lk
        sp = int(p.split)
        Q = RD(p.Q, B * Lq, H * 32, p.ldq, sp).view(B, Lq, -1).clone().requires_grad_(True)
        K = RD(p.K, B * Lk, H * 32, p.ldk, sp).view(B, Lk, -1).clone().requires_grad_(True)
        V = RD(p.V, B * Lk, H * 32, p.ldv, sp).view(B, Lk, -1).clone().requires_grad_(True)
        dO = RD(p.dO, B * Lq, H * 32, p.lddo, sp).view(B, Lq, -1)
        o, _ = self._attn_core(Q, K, V, p, B, H, Lq, Lk)
        gq, gk, gv = torch.autograd.grad(o, [Q, K, V], dO)
        parts = int(p.parts) or 7                     # 1 = delta (no observable output here), 2 = dK/dV, 4 = dQ
        if parts & 4:
            WR(p.dQ, B * Lq, H * 32, p.lddq, gq.reshape(B * Lq, -1), sp)
        if parts & 2:
            WR(p.dK, B * Lk, H * 32, p.lddk, gk.reshape(B * Lk, -1), sp)
            WR(p.dV, B * Lk, H * 32, p.lddv, gv.reshape(B * Lk, -1), sp)
        return 0

    # -- layer norm
    def detrb_layernorm_fwd(self, x, gamma, beta, y, y2, pos, S, mean, rstd, M, split, stream):
        S, M, sp = _v(S), _v(M), _v(split)
        xv = RD(x, M, 256, 256, sp)
        mu = xv.mean(-1, keepdim=True)
        var = ((xv - mu) ** 2).mean(-1, keepdim=True)
        rs = torch.rsqrt(var + 1e-5)
        o = (xv - mu) * rs * T(gamma, F32, 256) + T(beta, F32, 256)
        WR(y, M, 256, 256, o, sp)
        if _addr(y2):
            pv = RD(pos, S, 256, 256, sp)
            WR(y2, M, 256, 256, RND(o, sp) + pv[torch.arange(M) % S], sp)
        if _addr(mean):
            T(mean, F32, M)[:] = mu[:, 0]
        if _addr(rstd):
            T(rstd, F32, M)[:] = rs[:, 0]
        return 0

    def detrb_layernorm_bwd(self, dy, dy2, x, gamma, mean, rstd, dx, dx_drop, drop_p, seed, site, seed_ptr, dgamma, dbeta,
                            M, split, stream):
        M, drop_p, seed, site, sp = _v(M), _v(drop_p), _v(seed), _v(site), _v(split)
        d = RD(dy, M, 256, 256, sp)
        if _addr(dy2):
            d = d + RD(dy2, M, 256, 256, sp)
        xv = RD(x, M, 256, 256, sp)
        mu, rs = T(mean, F32, M)[:, None], T(rstd, F32, M)[:, None]
        g = T(gamma, F32, 256)
        xh = (xv - mu) * rs
        gv = d * g
        o = rs * (gv - gv.mean(-1, keepdim=True) - xh * (gv * xh).mean(-1, keepdim=True))
        WR(dx, M, 256, 256, o, sp)
        if _addr(dx_drop):
            od = RND(o, sp)
            if drop_p > 0:
                keep = keep_mask(range(M), range(256), drop_p, seed, site, seed_ptr)
                od = torch.where(keep, od / (1 - drop_p), torch.zeros_like(od))
            WR(dx_drop, M, 256, 256, od, sp)
        if _addr(dgamma):
            T(dgamma, F32, 256).add_((d * xh).sum(0))
            T(dbeta, F32, 256).add_(d.sum(0))
        return 0

    # -- elementwise
    def detrb_add_rowbcast(self, x, pos, out, M, S, d, split, stream):
        M, S, d, sp = _v(M), _v(S), _v(d), _v(split)
        xv = RD(x, M, d, d, sp)
        pv = RD(pos, S, d, d, sp)
        WR(out, M, d, d, xv + pv[torch.arange(M) % S], sp)
        return 0

    def detrb_add(self, a, b, out, n, stream):
        n = _v(n)
        v = T(a, _Act.dtype, n).to(F32)
        if _addr(b):
            v = v + T(b, _Act.dtype, n).to(F32)
        T(out, _Act.dtype, n)[:] = v.to(_Act.dtype)
        return 0

    def detrb_image_to_nhwc4(self, img, out, npix, stream):
        npix = _v(npix)
        o = M2(out, _Act.dtype, npix, 4, 4)
        o[:, :3] = M2(img, F32, npix, 3, 3).to(_Act.dtype)
        o[:, 3] = 0
        return 0

    def detrb_image_to_s2d16(self, img, out, B, H, W, pad_top, pad_left, HP, WP, split, stream):
        B, H, W, pt, pl, HP, WP = map(_v, (B, H, W, pad_top, pad_left, HP, WP))
        H2, W2 = (H + 1) // 2, (W + 1) // 2
        x = T(img, F32, B * H * W * 3).view(B, H, W, 3)
        xp = torch.zeros(B, 2 * H2, 2 * W2, 3)
        xp[:, :H, :W] = x
        o = torch.zeros(B, HP, WP, 16)
        for ry in range(2):
            for rx in range(2):
                o[:, pt:pt + H2, pl:pl + W2, (ry * 2 + rx) * 3:(ry * 2 + rx) * 3 + 3] = xp[:, ry::2, rx::2]
        WR(out, B * HP * WP, 16, 16, o.reshape(-1, 16), _v(split))
        return 0

    def detrb_f32_to_bf16(self, x, y, n, stream):
        n = _v(n)
        T(y, _Act.dtype, n)[:] = T(x, F32, n).to(_Act.dtype)
        return 0

    def detrb_colsum(self, x, ldx, M, N, scale, out, stream):
        ldx, M, N = _v(ldx), _v(M), _v(N)
        s = M2(x, _Act.dtype, M, N, ldx).to(F32).sum(0)
        if _addr(scale):
            s = s * T(scale, F32, N)
        T(out, F32, N).add_(s)
        return 0

    def detrb_maxpool_fwd(self, x, y, argmax, B, IH, IW, C, OH, OW, XH, XW, split, stream):
        B, IH, IW, C, OH, OW, XH, XW, sp = map(_v, (B, IH, IW, C, OH, OW, XH, XW, split))
        xv = RD(x, B * XH * XW, C, C, sp).view(B, XH, XW, C)[:, :IH, :IW]
        best = torch.full((B, OH, OW, C), -float("inf"))
        arg = torch.zeros((B, OH, OW, C), dtype=torch.uint8)
        oy, ox = torch.arange(OH), torch.arange(OW)
        for kh in range(3):
            for kw in range(3):
                iy, ix = oy * 2 - 1 + kh, ox * 2 - 1 + kw
                oky, okx = (iy >= 0) & (iy < IH), (ix >= 0) & (ix < IW)
                v = xv[:, iy.clamp(0, IH - 1)][:, :, ix.clamp(0, IW - 1)]
                ok = (oky[:, None] & okx[None, :])[None, :, :, None]
                v = torch.where(ok, v, torch.full_like(v, -float("inf")))
                upd = v > best
                best = torch.where(upd, v, best)
                arg = torch.where(upd, torch.full_like(arg, kh * 3 + kw), arg)
        arg = torch.where(best > 0, arg, torch.full_like(arg, 15))      # no gradient through the stem's ReLU
        WR(y, B * OH * OW, C, C, best.reshape(-1, C), sp)
        T(argmax, torch.uint8, B * OH * OW * C)[:] = arg.reshape(-1)
        return 0

    def detrb_maxpool_bwd(self, dy, argmax, dx, B, IH, IW, C, OH, OW, XH, XW, split, stream):
        B, IH, IW, C, OH, OW, XH, XW, sp = map(_v, (B, IH, IW, C, OH, OW, XH, XW, split))
        d = RD(dy, B * OH * OW, C, C, sp).view(B, OH, OW, C)
        arg = T(argmax, torch.uint8, B * OH * OW * C).view(B, OH, OW, C)
        out = torch.zeros(B, XH, XW, C)
        for oy in range(OH):
            for ox in range(OW):
                for kh in range(3):
                    for kw in range(3):
                        iy, ix = oy * 2 - 1 + kh, ox * 2 - 1 + kw
                        if 0 <= iy < IH and 0 <= ix < IW:
                            out[:, iy, ix] += d[:, oy, ox] * (arg[:, oy, ox] == kh * 3 + kw).to(F32)
        out[:, IH:] = 0
        out[:, :, IW:] = 0
        WR(dx, B * XH * XW, C, C, out.reshape(-1, C), sp)
        return 0

    # -- matcher / loss
    def detrb_matcher(self, logits, ldl, boxes, t_bbox, t_class, P, B, Q, C, fc, fb, fg, p_indices, t_indices, p_selector,
                      match, cost, status, stream):
        ldl, P, B, Q, C, fc, fb, fg = map(_v, (ldl, P, B, Q, C, fc, fb, fg))
        h = self.harness()
        lg = M2(logits, F32, P * Q, C, ldl).view(P, Q, C)
        bx = T(boxes, F32, P * Q * 4).view(P, Q, 4)
        tb = T(t_bbox, F32, B * 400).view(B, 100, 4)
        tc = T(t_class, torch.int64, B * 100).view(B, 100)
        pi = T(p_indices, torch.int64, P * Q).view(P, Q)
        ti = T(t_indices, torch.int64, P * Q).view(P, Q)
        ps = T(p_selector, torch.uint8, P * Q).view(P, Q)
        mt = T(match, torch.int32, P * Q).view(P, Q)
        st = T(status, torch.int32, P)
        co = T(cost, F32, P * Q * 100).view(P, Q, 100) if _addr(cost) else None
        for p in range(P):
            b = p % B
            n = int(tb[b, 0, 0])
            probs = torch.softmax(lg[p], -1).contiguous().numpy()
            pb = bx[p].contiguous().numpy()
            tbn = tb[b, 1:1 + n].contiguous().numpy()
            tcn = tc[b, 1:1 + n].contiguous().numpy()
            c = np.zeros((Q, max(n, 1)), np.float32)
            if n > 0:
                c = np.zeros((Q, n), np.float32)
                h.h_cost_matrix(pb.ctypes.data_as(ctypes.c_void_p), probs.ctypes.data_as(ctypes.c_void_p), Q, C,
                                tbn.ctypes.data_as(ctypes.c_void_p), tcn.ctypes.data_as(ctypes.c_void_p), n,
                                ctypes.c_float(fc), ctypes.c_float(fb), ctypes.c_float(fg), c.ctypes.data_as(ctypes.c_void_p))
            if co is not None and n > 0:
                co[p, :, :n] = torch.from_numpy(c)
            r4c = np.full(Q, -1, np.int32)
            rc = 0
            if n > 0:
                costT = np.ascontiguousarray(c.T)
                rc = h.lsap_warp_model(costT.ctypes.data_as(ctypes.c_void_p), n, Q, r4c.ctypes.data_as(ctypes.c_void_p))
            st[p] = rc
            mt[p] = torch.from_numpy(r4c)
            ps[p] = torch.from_numpy((r4c >= 0).astype(np.uint8))
            qs = np.nonzero(r4c >= 0)[0]
            pi[p] = -1
            ti[p] = -1
            pi[p, :len(qs)] = torch.from_numpy(qs.astype(np.int64))
            ti[p, :len(qs)] = torch.from_numpy(r4c[qs].astype(np.int64))
        return 0

    def detrb_set_loss(self, logits, ldl, boxes, t_bbox, t_class, match, L, B, Q, C, bg, normalisers, loss_scale, sums, losses,
                       total, d_logits, ld_dl, d_boxpre, ld_db, status, split, stream):
        ldl, L, B, Q, C, bg, loss_scale, ld_dl, ld_db, sp = map(_v, (ldl, L, B, Q, C, bg, loss_scale, ld_dl, ld_db, split))
        lg = M2(logits, F32, L * B * Q, C, ldl).clone().view(L, B, Q, C).requires_grad_(True)
        bx = T(boxes, F32, L * B * Q * 4).clone().view(L, B, Q, 4).requires_grad_(True)
        tb = T(t_bbox, F32, B * 400).view(B, 100, 4)
        tc = T(t_class, torch.int64, B * 100).view(B, 100)
        mt = T(match, torch.int32, L * B * Q).view(L, B, Q).long()
        ns = tb[:, 0, 0].clamp(0, 99)
        if _addr(normalisers):
            nrm = T(normalisers, F32, 2)
            n_matched, sum_w = float(nrm[0]), float(nrm[1])
        else:
            n_matched = float(ns.sum())
            sum_w = 0.1 * (B * Q - n_matched) + n_matched
        out = T(losses, F32, L * 6).view(L, 6)
        tot = 0
        for l in range(L):
            m = mt[l]
            matched = m >= 0
            cls = torch.where(matched, torch.gather(tc, 1, (m.clamp(min=0) + 1)), torch.full_like(m, bg))
            w = torch.where(matched, torch.ones(B, Q), torch.full((B, Q), 0.1))
            ce = F.cross_entropy(lg[l].reshape(B * Q, C), cls.reshape(-1), reduction="none").view(B, Q)
            label = (ce * w).sum() / sum_w
            am = lg[l].argmax(-1)
            tbm = torch.gather(tb, 1, (m.clamp(min=0) + 1)[..., None].expand(B, Q, 4))
            pbm = bx[l]
            l1 = ((pbm - tbm).abs().sum(-1) * matched).sum() / n_matched

            def xyxy(b_):
                return torch.cat([b_[..., :2] - b_[..., 2:] / 2, b_[..., :2] + b_[..., 2:] / 2], -1).clamp(0, 1)
            pa, ta = xyxy(pbm), xyxy(tbm)
            inter = (torch.minimum(pa[..., 2:], ta[..., 2:]) - torch.maximum(pa[..., :2], ta[..., :2])).clamp(min=0)
            inter = inter[..., 0] * inter[..., 1]
            ap = (pa[..., 2] - pa[..., 0]) * (pa[..., 3] - pa[..., 1])
            at = (ta[..., 2] - ta[..., 0]) * (ta[..., 3] - ta[..., 1])
            uni = ap + at - inter
            cwh = (torch.maximum(pa[..., 2:], ta[..., 2:]) - torch.minimum(pa[..., :2], ta[..., :2])).clamp(min=0)
            area = cwh[..., 0] * cwh[..., 1]
            giou = inter / uni - (area - uni) / area
            gl = torch.where(matched, 1 - giou, torch.zeros_like(giou)).sum() / n_matched
            out[l, 0] = label.detach()
            out[l, 1] = ((am == bg) & ~matched).sum() / (~matched).sum()
            out[l, 2] = ((am != bg) & matched).sum() / matched.sum()
            out[l, 3] = ((am == cls) & matched).sum() / matched.sum()
            out[l, 4] = gl.detach()
            out[l, 5] = l1.detach()
            tot = tot + label + 2 * gl + 5 * l1
        tot = tot * loss_scale
        T(total, F32, 1)[0] = tot.detach()
        if _addr(status) and int(T(status, torch.int32, L * B).abs().sum()) != 0:      # NaN / -inf cost matrix: poisoned result
            T(total, F32, 1)[0] = float("nan")
            out[:] = float("nan")
        if _addr(d_logits):
            gl_, gb_ = torch.autograd.grad(tot, [lg, bx])
            dl = torch.zeros(L * B * Q, ld_dl)
            dl[:, :C] = gl_.reshape(-1, C)
            WR(d_logits, L * B * Q, ld_dl, ld_dl, dl, sp)
            db = torch.zeros(L * B * Q, ld_db)
            bxd = bx.detach()
            db[:, :4] = (gb_ * bxd * (1 - bxd)).reshape(-1, 4)
            WR(d_boxpre, L * B * Q, ld_db, ld_db, db, sp)
        return 0

    # -- optimizer
    def detrb_adam_clipnorm(self, params, grads, m, v, table, lr_group, lrs, enabled, Tn, total, clipnorm, beta1, beta2, eps,
                            steps, norms, stream):
        Tn, total, clipnorm, beta1, beta2, eps = map(_v, (Tn, total, clipnorm, beta1, beta2, eps))
        P_, G_, M_, V_ = (T(x, F32, total) for x in (params, grads, m, v))
        tab = T(table, torch.int64, 2 * Tn).view(Tn, 2)
        grp = T(lr_group, torch.int32, Tn)
        lr = T(lrs, F32, 8)
        en = T(enabled, torch.uint8, 8)
        st = T(steps, torch.int32, 8)
        nr = T(norms, F32, Tn)
        st += en.to(torch.int32)
        for t in range(Tn):
            o, n = int(tab[t, 0]), int(tab[t, 1])
            g = G_[o:o + n]
            nr[t] = (g * g).sum()
            k = int(grp[t])
            if not en[k]:
                continue
            norm = float(nr[t].sqrt())
            g = g * (clipnorm / norm) if (clipnorm > 0 and norm > clipnorm) else g
            step = float(st[k])
            M_[o:o + n] = beta1 * M_[o:o + n] + (1 - beta1) * g
            V_[o:o + n] = beta2 * V_[o:o + n] + (1 - beta2) * g * g
            lr_t = float(lr[k]) * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
            P_[o:o + n] -= lr_t * M_[o:o + n] / (V_[o:o + n].sqrt() + eps)
        return 0

    def detrb_adam_clipnorm_chunked(self, params, grads, m, v, chunks, nchunks, lr_group, lrs, enabled, Tn, clipnorm, beta1, beta2, eps,
                                    steps, norms, prologue, stream):
        """one optimizer step issued over a chunk range holding whole variables (see include/detrb.h)"""
        nchunks, Tn, clipnorm, beta1, beta2, eps, prologue = map(_v, (nchunks, Tn, clipnorm, beta1, beta2, eps, prologue))
        ch = T(chunks, torch.int32, 3 * nchunks).view(nchunks, 3)
        grp, lr, en = T(lr_group, torch.int32, Tn), T(lrs, F32, 8), T(enabled, torch.uint8, 8)
        st, nr = T(steps, torch.int32, 8), T(norms, F32, Tn)
        if prologue:
            st += en.to(torch.int32)
            nr.zero_()
        end = int((ch[:, 1] + ch[:, 2]).max())
        P_, G_, M_, V_ = (T(x, F32, end) for x in (params, grads, m, v))
        for t, o, n in ch.tolist():
            g = G_[o:o + n]
            nr[t] += (g * g).sum()
        for t, o, n in ch.tolist():
            k = int(grp[t])
            if not en[k]:
                continue
            norm = float(nr[t].sqrt())
            g = G_[o:o + n]
            g = g * (clipnorm / norm) if (clipnorm > 0 and norm > clipnorm) else g
            step = float(st[k])
            M_[o:o + n] = beta1 * M_[o:o + n] + (1 - beta1) * g
            V_[o:o + n] = beta2 * V_[o:o + n] + (1 - beta2) * g * g
            lr_t = float(lr[k]) * np.sqrt(1 - beta2 ** step) / (1 - beta1 ** step)
            P_[o:o + n] -= lr_t * M_[o:o + n] / (V_[o:o + n].sqrt() + eps)
        return 0

    def detrb_prep_weight(self, master, fold, N, taps, Cin, Wf, ldf, Wd, ldd, stream):
        N, taps, Cin, ldf, ldd = map(_v, (N, taps, Cin, ldf, ldd))
        w = T(master, F32, N * taps * Cin).view(N, taps, Cin)
        if _addr(fold):
            w = w * T(fold, F32, N)[:, None, None]
        wb = w.to(_Act.dtype)
        if _addr(Wf):
            M2(Wf, _Act.dtype, N, taps * Cin, ldf)[:] = wb.reshape(N, taps * Cin)
        if _addr(Wd):
            t = T(Wd, _Act.dtype, Cin * taps * ldd).view(Cin, taps, ldd)
            t[:, :, :N] = wb.permute(2, 1, 0)
        return 0

    def detrb_prep_weights_multi(self, descs, nslots, total_tiles, wsplit, stream):
        n, ws = _v(nslots), _v(wsplit)
        arr = (_lib.PrepDesc * n).from_address(_addr(descs))
        for d in arr:
            w = T(d.master, F32, d.N * d.taps * d.Cin).view(d.N, d.taps, d.Cin)
            if d.fold:
                w = w * T(d.fold, F32, d.N)[:, None, None]
            if d.Wf:
                WR(d.Wf, d.N, d.taps * d.Cin, d.ldf, w.reshape(d.N, d.taps * d.Cin), ws)
            if d.Wd:
                WR(d.Wd, d.Cin * d.taps, d.N, d.ldd, w.permute(2, 1, 0).reshape(d.Cin * d.taps, d.N), ws)
        return 0

    def detrb_accumulate(self, acc, g, n, zero_first, stream):
        n = _v(n)
        a, b = T(acc, F32, n), T(g, F32, n)
        if _v(zero_first):
            a.zero_()
        a.add_(b)
        return 0

    # ---- rows either side of the train step (csrc/pipeline.cu)
    @staticmethod
    def _lut_apply(img, lut, swap, npix):
        x = T(img, torch.uint8, npix * 3).view(npix, 3).long()
        l = T(lut, F32, 768).view(3, 256)
        if _v(swap):
            x = x.flip(-1)
        return torch.stack([l[c][x[:, c]] for c in range(3)], -1)

    def detrb_normalize_u8(self, img, lut, swap, out, npix, stream):
        npix = _v(npix)
        T(out, F32, npix * 3)[:] = self._lut_apply(img, lut, swap, npix).reshape(-1)
        return 0

    def detrb_image_u8_to_s2d16(self, img, lut, swap, out, B, H, W, pad_top, pad_left, HP, WP, split, stream):
        B, H, W = _v(B), _v(H), _v(W)
        x = self._lut_apply(img, lut, swap, B * H * W).contiguous()
        return self.detrb_image_to_s2d16(ctypes.c_void_p(x.data_ptr()), out, B, H, W, pad_top, pad_left, HP, WP, split, stream)

    def detrb_resize_affine_u8(self, src, src_off, src_hw, inv, zero_border, out, B, H, W, stream):
        from oracle.resize_oracle import resize_affine_u8
        B, H, W = _v(B), _v(H), _v(W)
        off, hw = T(src_off, torch.int64, B).numpy(), T(src_hw, torch.int32, 2 * B).numpy()
        frames = [T(ctypes.c_void_p(_addr(src) + int(off[b])), torch.uint8, int(hw[2 * b]) * int(hw[2 * b + 1]) * 3).view(int(hw[2 * b]), int(hw[2 * b + 1]), 3).numpy()
                  for b in range(B)]
        res = resize_affine_u8(frames, T(inv, F32, 4 * B).view(B, 4).numpy(), T(zero_border, torch.uint8, B).numpy(), H, W)
        T(out, torch.uint8, B * H * W * 3)[:] = torch.from_numpy(res).reshape(-1)
        return 0

    def detrb_postprocess(self, logits, ldl, boxes, B, Q, C, bg, fmt, out_boxes, out_labels, out_scores, out_query, out_count,
                          stream):
        B, Q, C, ldl, bg, fmt = _v(B), _v(Q), _v(C), _v(ldl), _v(bg), _v(fmt)
        lg = M2(logits, F32, B * Q, C, ldl).view(B, Q, C)
        bx = T(boxes, F32, B * Q * 4).view(B, Q, 4)
        ob, ol = T(out_boxes, F32, B * Q * 4).view(B, Q, 4), T(out_labels, torch.int64, B * Q).view(B, Q)
        os_, oc = T(out_scores, F32, B * Q).view(B, Q), T(out_count, torch.int32, B)
        oq = T(out_query, torch.int32, B * Q)
        for b in range(B):
            e = torch.exp(lg[b] - lg[b].max(-1, keepdim=True).values)
            label = torch.from_numpy(np.argmax(e.numpy(), -1))
            score = e.max(-1).values / e.sum(-1)
            keep = torch.nonzero(label != bg).squeeze(-1)
            k = keep.numel()
            r = bx[b][keep]
            if fmt != 0:
                c = torch.cat([r[:, :2] - r[:, 2:] * 0.5, r[:, :2] + r[:, 2:] * 0.5], -1).clamp(0.0, 1.0)
                r = c if fmt == 1 else c[:, [1, 0, 3, 2]]
            ob[b, :k], ol[b, :k], os_[b, :k], oc[b] = r, label[keep], score[keep], k
            if oq is not None:
                oq.view(B, Q)[b, :k] = keep.int()
        return 0

    def detrb_map_match(self, pred_boxes, pred_labels, pred_scores, pred_count, B, Q, t_boxes, t_labels, t_count, NT, t_wire, thresholds, T,
                        num_classes, rank, tp, gt_count, stream):
        B, Q, NT, t_wire, Tn, ncls = map(_v, (B, Q, NT, t_wire, T, num_classes))
        pb = T_(pred_boxes, F32, B * Q * 4).view(B, Q, 4).numpy()
        pl = T_(pred_labels, torch.int64, B * Q).view(B, Q).numpy()
        ps = T_(pred_scores, F32, B * Q).view(B, Q).numpy()
        pc = T_(pred_count, torch.int32, B).numpy()
        thr = T_(thresholds, torch.float64, Tn).numpy()
        rk = T_(rank, torch.int32, B * Q).view(B, Q)
        out = T_(tp, torch.uint8, B * Tn * Q).view(B, Tn, Q)
        gc = T_(gt_count, torch.int32, ncls) if _addr(gt_count) else None
        f = np.float32
        for b in range(B):
            k = int(min(max(pc[b], 0), Q))
            if t_wire:
                tb_ = T_(t_boxes, F32, B * NT * 4).view(B, NT, 4).numpy()[b]
                n = int(min(max(tb_[0, 0], 0), NT - 1))
                c = tb_[1:1 + n]
                xy = np.clip(np.concatenate([c[:, :2] - c[:, 2:] * f(0.5), c[:, :2] + c[:, 2:] * f(0.5)], -1), f(0), f(1)).astype(f)
                gtb = xy[:, [1, 0, 3, 2]]
                gtc = T_(t_labels, torch.int64, B * NT).view(B, NT).numpy()[b, 1:1 + n]
            else:
                n = int(min(max(T_(t_count, torch.int32, B).numpy()[b], 0), NT))
                gtb = T_(t_boxes, F32, B * NT * 4).view(B, NT, 4).numpy()[b, :n]
                gtc = T_(t_labels, torch.int64, B * NT).view(B, NT).numpy()[b, :n]
            if gc is not None:
                for cc in gtc:
                    if 0 <= cc < ncls:
                        gc[int(cc)] += 1
            order = sorted(range(k), key=lambda i: -float(ps[b, i]))
            rk[b] = -1
            for r, i in enumerate(order):
                rk[b, i] = r
            ga = (gtb[:, 2] - gtb[:, 0]) * (gtb[:, 3] - gtb[:, 1])
            out[b] = 0
            with np.errstate(invalid="ignore", divide="ignore"):
                for t in range(Tn):
                    used = [False] * n
                    for i in order:
                        q = pb[b, i]
                        qa = (q[2] - q[0]) * (q[3] - q[1])
                        best, bj = float(thr[t]), -1
                        for j in range(n):
                            if used[j] or gtc[j] != pl[b, i]:
                                continue
                            g = gtb[j]
                            inter = max(min(g[3], q[3]) - max(g[1], q[1]), f(0)) * max(min(g[2], q[2]) - max(g[0], q[0]), f(0))
                            v = float(f(inter) / f(f(ga[j] + qa) - f(inter)))
                            if v > best:
                                best, bj = v, j
                        if bj >= 0:
                            used[bj] = True
                            out[b, t, i] = 1
        return 0

    def detrb_attn_dropout_mask(self, out, M, N, drop_p, seed, site, seed_ptr, stream):
        M, N = _v(M), _v(N)
        T(out, torch.uint8, M * N)[:] = attn_keep_mask(range(M), range(N), _v(drop_p), _v(seed), _v(site), seed_ptr).reshape(-1).to(torch.uint8)
        return 0

    def detrb_dropout_mask(self, out, M, N, drop_p, seed, site, seed_ptr, stream):
        M, N = _v(M), _v(N)
        T(out, torch.uint8, M * N)[:] = keep_mask(range(M), range(N), _v(drop_p), _v(seed), _v(site), seed_ptr).reshape(-1).to(torch.uint8)
        return 0


def install():
    """Route detr_tensorflow_b200's C-ABI calls to the emulator (CPU tensors stand in for device memory)."""
    fake = FakeLib()
    _lib._lib = fake
    _lib._EMULATED = True
    return fake


def uninstall():
    _lib._lib = None
    _lib._EMULATED = False
