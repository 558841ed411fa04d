"""Thin Python wrappers over the C ABI (include/detrb.h).  Arguments are torch tensors (device memory owners)
and plain ints; every call is enqueued on torch's current CUDA stream.  No arithmetic happens here."""
from __future__ import annotations

import ctypes
from ctypes import c_float, c_int, c_int64, c_uint32, c_uint64, c_void_p, byref

import torch

from . import _lib
from ._lib import AttnBwdParams, AttnFwdParams, IgemmParams, WgradParams, check


def _stream():
    if torch.cuda.is_available():
        return c_void_p(torch.cuda.current_stream().cuda_stream)
    return c_void_p(0)


def ptr(t, offset=0):
    """device pointer of tensor `t` (+ element offset); None -> NULL"""
    if t is None:
        return None
    return c_void_p(t.data_ptr() + offset * t.element_size())


def plain_geom(M, K):
    return dict(batch=1, IH=1, IW=M, Cin=K, OH=1, OW=M, KH=1, KW=1, stride=1, pad=0, mode=0)


def igemm(A, W, M, N, K, lda, ldw, geom, *, bias=None, residual=None, ldr=0, mask=None, ldm=0, mask_scale=1.0,
          relu=False, sigmoid=False, drop_p=0.0, seed=0, site=0, seed_ptr=None, C=None, ldc=0, Cf=None, ldcf=0,
          out_stride=1, SH=0, SW=0, accumulate=False, a_kb_rows=0, force_tc=None, split=0, wsplit=0, mask_bits=None, ldmb=0, out_bits=None, ldob=0, scratch=None):
    p = IgemmParams()
    p.A, p.W = ptr(A), ptr(W)
    p.M, p.N, p.K, p.lda, p.ldw = M, N, K, lda, ldw
    for k, v in geom.items():
        setattr(p, k, v)
    p.bias, p.residual, p.ldr, p.mask, p.ldm, p.mask_scale = ptr(bias), ptr(residual), ldr, ptr(mask), ldm, mask_scale
    p.relu, p.sigmoid, p.drop_p, p.seed, p.site, p.seed_ptr = int(relu), int(sigmoid), drop_p, seed, site, ptr(seed_ptr)
    p.C, p.ldc, p.Cf, p.ldcf = ptr(C), ldc, ptr(Cf), ldcf
    p.out_stride, p.SH, p.SW, p.accumulate, p.a_kb_rows = out_stride, SH, SW, int(accumulate), a_kb_rows
    p.split, p.wsplit = split, wsplit
    p.mask_bits, p.ldmb, p.out_bits, p.ldob = ptr(mask_bits), ldmb, ptr(out_bits), ldob
    p.scratch = ptr(scratch)
    if force_tc is not None:
        check(_lib.lib().detrb_gemm_tc_force(byref(p), c_int(force_tc), _stream()))
    else:
        check(_lib.lib().detrb_igemm(byref(p), _stream()))


OPTIONS = {"pdl": 0, "tc": 1, "tc_conv": 2, "tc_tma_epilogue": 3, "tc_persistent": 4, "tc_stream": 5, "tc_halo": 6, "tc_pair": 7,
           "tc_wgrad": 8, "tc_attn": 9}          # DETRB_OPT_* of include/detrb.h


class Handle:
    """detrb_handle_t (include/detrb.h, "Handles"): the device a step runs on + the kernel-policy switches, one per GPU / rank.
    `bind()` makes it current for the calling thread (cudaSetDevice + the thread-local switches)."""

    def __init__(self, device_index):
        L = _lib.lib()
        self._h = c_void_p()
        check(L.detrb_create(c_int(int(device_index)), byref(self._h)))

    @property
    def device_index(self):
        return int(_lib.lib().detrb_handle_device(self._h))

    def set(self, option, value):
        check(_lib.lib().detrb_handle_set(self._h, c_int(OPTIONS[option]), c_int(int(value))))

    def get(self, option):
        v = c_int()
        check(_lib.lib().detrb_handle_get(self._h, c_int(OPTIONS[option]), byref(v)))
        return v.value

    def bind(self):
        check(_lib.lib().detrb_bind(self._h))

    def close(self):
        if self._h:
            _lib.lib().detrb_destroy(self._h)
            self._h = c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def set_tc(enable):
    """route plain GEMMs through the tcgen05/TMA/TMEM kernel (gemm_tc.cu); returns the previous setting"""
    return _lib.lib().detrb_set_tc(c_int(int(enable)))


def set_tc_wgrad(enable):
    return _lib.lib().detrb_set_tc_wgrad(c_int(int(enable)))


def set_wgrad_tile(mode):
    """tile of the tcgen05 weight-gradient kernel: 0 policy, 1 128x128, 2 128x256, 3 256x128, 4 256x256; returns the previous mode"""
    return int(_lib.lib().detrb_set_wgrad_tile(c_int(int(mode))))


def set_tc_attn(enable):
    """attention forward on the tcgen05 kernel (attention_tc.cu); returns the previous setting"""
    return _lib.lib().detrb_set_tc_attn(c_int(int(enable)))


def set_pdl(enable):
    return _lib.lib().detrb_set_pdl(c_int(int(enable)))


def set_tc_persistent(enable):
    return _lib.lib().detrb_set_tc_persistent(c_int(int(enable)))


def set_tc_stream(enable):
    return _lib.lib().detrb_set_tc_stream(c_int(int(enable)))


def set_tc_pair(mode):
    """CTA-pair persistent GEMM (cta_group::2): 0 off, 1 256-wide tiles, 2 128- and 256-wide tiles, -1 environment; returns the old mode."""
    return int(_lib.lib().detrb_set_tc_pair(c_int(mode)))


def set_tc_halo(enable):
    return _lib.lib().detrb_set_tc_halo(c_int(int(enable)))


def set_tc_conv(enable):
    return _lib.lib().detrb_set_tc_conv(c_int(int(enable)))


def wgrad(A, lda, dY, ldy, M, N, K, geom, dW, ldw, *, rowscale=None, dbias=None, a_kb_rows=0, k_mask=False, force_tc=False, split=0):
    p = WgradParams()
    p.A, p.lda, p.dY, p.ldy, p.M, p.N, p.K = ptr(A), lda, ptr(dY), ldy, M, N, K
    for k, v in geom.items():
        if k != "mode":
            setattr(p, k, v)
    p.rowscale, p.dW, p.ldw, p.dbias = ptr(rowscale), ptr(dW), ldw, ptr(dbias)
    p.a_kb_rows, p.k_mask, p.split = a_kb_rows, int(k_mask), split
    if force_tc:
        check(_lib.lib().detrb_wgrad_tc_force(byref(p), _stream()))
    else:
        check(_lib.lib().detrb_wgrad(byref(p), _stream()))


def attn_fwd(Q, K, V, ldq, ldk, ldv, O, ldo, lse, B, H, Lq, Lk, scale, drop_p=0.0, seed=0, site=0, seed_ptr=None, split=0):
    p = AttnFwdParams()
    p.Q, p.K, p.V, p.ldq, p.ldk, p.ldv = ptr(Q), ptr(K), ptr(V), ldq, ldk, ldv
    p.O, p.ldo, p.lse, p.B, p.H, p.Lq, p.Lk = ptr(O), ldo, ptr(lse), B, H, Lq, Lk
    p.scale, p.drop_p, p.seed, p.site, p.seed_ptr, p.split = scale, drop_p, seed, site, ptr(seed_ptr), split
    check(_lib.lib().detrb_attn_fwd(byref(p), _stream()))


def attn_bwd(Q, K, V, O, dO, ldq, ldk, ldv, ldo, lddo, lse, delta, dQ, dK, dV, lddq, lddk, lddv, B, H, Lq, Lk,
             scale, drop_p=0.0, seed=0, site=0, seed_ptr=None, split=0, parts=0):
    p = AttnBwdParams()
    p.Q, p.K, p.V, p.O, p.dO = ptr(Q), ptr(K), ptr(V), ptr(O), ptr(dO)
    p.ldq, p.ldk, p.ldv, p.ldo, p.lddo = ldq, ldk, ldv, ldo, lddo
    p.lse, p.delta, p.dQ, p.dK, p.dV = ptr(lse), ptr(delta), ptr(dQ), ptr(dK), ptr(dV)
    p.lddq, p.lddk, p.lddv, p.B, p.H, p.Lq, p.Lk = lddq, lddk, lddv, B, H, Lq, Lk
    p.scale, p.drop_p, p.seed, p.site, p.seed_ptr, p.split, p.parts = scale, drop_p, seed, site, ptr(seed_ptr), split, parts
    check(_lib.lib().detrb_attn_bwd(byref(p), _stream()))


def layernorm_fwd(x, gamma, beta, y, y2, pos, S, mean, rstd, M, split=0):
    check(_lib.lib().detrb_layernorm_fwd(ptr(x), ptr(gamma), ptr(beta), ptr(y), ptr(y2), ptr(pos), c_int(S),
                                         ptr(mean), ptr(rstd), c_int(M), c_int64(split), _stream()))


def layernorm_bwd(dy, dy2, x, gamma, mean, rstd, dx, dx_drop, drop_p, seed, site, seed_ptr, dgamma, dbeta, M, split=0):
    check(_lib.lib().detrb_layernorm_bwd(ptr(dy), ptr(dy2), ptr(x), ptr(gamma), ptr(mean), ptr(rstd), ptr(dx),
                                         ptr(dx_drop), c_float(drop_p), c_uint64(seed), c_uint32(site), ptr(seed_ptr),
                                         ptr(dgamma), ptr(dbeta), c_int(M), c_int64(split), _stream()))


def add_rowbcast(x, pos, out, M, S, d, split=0):
    check(_lib.lib().detrb_add_rowbcast(ptr(x), ptr(pos), ptr(out), c_int(M), c_int(S), c_int(d), c_int64(split), _stream()))


def add(a, b, out, n):
    check(_lib.lib().detrb_add(ptr(a), ptr(b), ptr(out), c_int64(n), _stream()))


def image_to_nhwc4(img, out, npix):
    check(_lib.lib().detrb_image_to_nhwc4(ptr(img), ptr(out), c_int64(npix), _stream()))


def image_to_s2d16(img, out, B, H, W, pad_top=0, pad_left=0, HP=None, WP=None, split=0):
    HP, WP = HP or (H + 1) // 2, WP or (W + 1) // 2
    check(_lib.lib().detrb_image_to_s2d16(ptr(img), ptr(out), c_int(B), c_int(H), c_int(W), c_int(pad_top), c_int(pad_left),
                                          c_int(HP), c_int(WP), c_int64(split), _stream()))


def f32_to_bf16(x, y, n):
    check(_lib.lib().detrb_f32_to_bf16(ptr(x), ptr(y), c_int64(n), _stream()))


def colsum(x, ldx, M, N, scale, out):
    check(_lib.lib().detrb_colsum(ptr(x), c_int(ldx), c_int(M), c_int(N), ptr(scale), ptr(out), _stream()))


def maxpool_fwd(x, y, argmax, B, IH, IW, C, OH, OW, XH=None, XW=None, split=0):
    check(_lib.lib().detrb_maxpool_fwd(ptr(x), ptr(y), ptr(argmax), c_int(B), c_int(IH), c_int(IW), c_int(C),
                                       c_int(OH), c_int(OW), c_int(XH or IH), c_int(XW or IW), c_int64(split), _stream()))


def maxpool_bwd(dy, argmax, dx, B, IH, IW, C, OH, OW, XH=None, XW=None, split=0):
    check(_lib.lib().detrb_maxpool_bwd(ptr(dy), ptr(argmax), ptr(dx), c_int(B), c_int(IH), c_int(IW), c_int(C),
                                       c_int(OH), c_int(OW), c_int(XH or IH), c_int(XW or IW), c_int64(split), _stream()))


def matcher(logits, ldl, boxes, t_bbox, t_class, P, B, Q, C, p_indices, t_indices, p_selector, match, cost, status,
            fcost_class=1.0, fcost_bbox=5.0, fcost_giou=2.0):
    check(_lib.lib().detrb_matcher(ptr(logits), c_int(ldl), ptr(boxes), ptr(t_bbox), ptr(t_class), c_int(P), c_int(B),
                                   c_int(Q), c_int(C), c_float(fcost_class), c_float(fcost_bbox), c_float(fcost_giou),
                                   ptr(p_indices), ptr(t_indices), ptr(p_selector), ptr(match), ptr(cost), ptr(status),
                                   _stream()))


def set_loss(logits, ldl, boxes, t_bbox, t_class, match, L, B, Q, C, background_class, normalisers, loss_scale,
             sums, losses, total, d_logits, ld_dl, d_boxpre, ld_db, status=None, split=0):
    check(_lib.lib().detrb_set_loss(ptr(logits), c_int(ldl), ptr(boxes), ptr(t_bbox), ptr(t_class), ptr(match),
                                    c_int(L), c_int(B), c_int(Q), c_int(C), c_int(background_class), ptr(normalisers),
                                    c_float(loss_scale), ptr(sums), ptr(losses), ptr(total), ptr(d_logits), c_int(ld_dl),
                                    ptr(d_boxpre), c_int(ld_db), ptr(status), c_int64(split), _stream()))


def adam_clipnorm(params, grads, m, v, table, lr_group, lrs, group_enabled, T, total, clipnorm, steps, norms,
                  beta1=0.9, beta2=0.999, eps=1e-7):
    check(_lib.lib().detrb_adam_clipnorm(ptr(params), ptr(grads), ptr(m), ptr(v), ptr(table), ptr(lr_group), ptr(lrs),
                                         ptr(group_enabled), c_int(T), c_int64(total), c_float(clipnorm), c_float(beta1),
                                         c_float(beta2), c_float(eps), ptr(steps), ptr(norms), _stream()))


def prep_weight(master, fold, N, taps, Cin, Wf, ldf, Wd, ldd):
    check(_lib.lib().detrb_prep_weight(ptr(master), ptr(fold), c_int(N), c_int(taps), c_int(Cin), ptr(Wf), c_int(ldf),
                                       ptr(Wd), c_int(ldd), _stream()))


def dropout_mask(out, M, N, drop_p, seed, site, seed_ptr=None):
    check(_lib.lib().detrb_dropout_mask(ptr(out), c_int(M), c_int(N), c_float(drop_p), c_uint64(seed), c_uint32(site),
                                        ptr(seed_ptr), _stream()))


def attn_dropout_mask(out, M, N, drop_p, seed, site, seed_ptr=None):
    check(_lib.lib().detrb_attn_dropout_mask(ptr(out), c_int(M), c_int(N), c_float(drop_p), c_uint64(seed), c_uint32(site),
                                             ptr(seed_ptr), _stream()))


def prep_weights_multi(descs_dev, nslots, total_tiles, wsplit=0):
    check(_lib.lib().detrb_prep_weights_multi(ptr(descs_dev), c_int(nslots), c_int(total_tiles), c_int64(wsplit), _stream()))


def accumulate(acc, g, n, zero_first):
    """acc[:n] = (0 if zero_first else acc[:n]) + g[:n]  (fp32; optimizers.py:150-157)"""
    check(_lib.lib().detrb_accumulate(ptr(acc), ptr(g), c_int64(n), c_int(int(zero_first)), _stream()))


def adam_clipnorm_chunked(params, grads, m, v, chunks, nchunks, lr_group, lrs, group_enabled, T, clipnorm, steps, norms,
                          beta1=0.9, beta2=0.999, eps=1e-7, first_chunk=0, prologue=True):
    """chunks [first_chunk, first_chunk + nchunks) of the [*, 3] int32 chunk table"""
    check(_lib.lib().detrb_adam_clipnorm_chunked(ptr(params), ptr(grads), ptr(m), ptr(v), ptr(chunks, 3 * first_chunk), c_int(nchunks),
                                                 ptr(lr_group), ptr(lrs), ptr(group_enabled), c_int(T), c_float(clipnorm), c_float(beta1),
                                                 c_float(beta2), c_float(eps), ptr(steps), ptr(norms), c_int(int(prologue)), _stream()))


def normalize_u8(img_u8, lut, swap_rb, out_f32, npix):
    check(_lib.lib().detrb_normalize_u8(ptr(img_u8), ptr(lut), c_int(int(swap_rb)), ptr(out_f32), c_int64(npix), _stream()))


def image_u8_to_s2d16(img_u8, lut, swap_rb, out, B, H, W, pad_top=0, pad_left=0, HP=None, WP=None, split=0):
    HP, WP = HP or (H + 1) // 2, WP or (W + 1) // 2
    check(_lib.lib().detrb_image_u8_to_s2d16(ptr(img_u8), ptr(lut), c_int(int(swap_rb)), ptr(out), c_int(B), c_int(H), c_int(W),
                                             c_int(pad_top), c_int(pad_left), c_int(HP), c_int(WP), c_int64(split), _stream()))


def postprocess(logits, ldl, boxes, B, Q, C, background_class, bbox_format, out_boxes, out_labels, out_scores, out_query,
                out_count):
    check(_lib.lib().detrb_postprocess(ptr(logits), c_int(ldl), ptr(boxes), c_int(B), c_int(Q), c_int(C),
                                       c_int(int(background_class)), c_int(int(bbox_format)), ptr(out_boxes), ptr(out_labels),
                                       ptr(out_scores), ptr(out_query), ptr(out_count), _stream()))


def map_match(pred_boxes, pred_labels, pred_scores, pred_count, B, Q, t_boxes, t_labels, t_count, NT, t_wire, thresholds, T, num_classes,
              rank, tp, gt_count):
    check(_lib.lib().detrb_map_match(ptr(pred_boxes), ptr(pred_labels), ptr(pred_scores), ptr(pred_count), c_int(B), c_int(Q), ptr(t_boxes),
                                     ptr(t_labels), ptr(t_count), c_int(NT), c_int(int(t_wire)), ptr(thresholds), c_int(T), c_int(num_classes),
                                     ptr(rank), ptr(tp), ptr(gt_count), _stream()))


def resize_affine_u8(src, src_off, src_hw, inv, zero_border, out, B, H, W):
    check(_lib.lib().detrb_resize_affine_u8(ptr(src), ptr(src_off), ptr(src_hw), ptr(inv), ptr(zero_border), ptr(out), c_int(B), c_int(H),
                                            c_int(W), _stream()))
