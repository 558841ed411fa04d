"""ctypes binding of libdetrb.so (C ABI declared in include/detrb.h).

The shared object is built in-tree by ``build()`` (nvcc, sm_100a only) and loaded lazily.  There is no
fallback: if the library is missing or the device is not a B200-class GPU every entry point raises.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import (POINTER, Structure, c_char_p, c_float, c_int, c_int32, c_int64, c_uint8, c_uint16, c_uint32,
                    c_uint64, c_void_p)

_HERE = os.path.dirname(os.path.abspath(__file__))
_CSRC = os.path.join(_HERE, "csrc")
# DETRB_SO: developer override (e.g. a -DDETRB_TRACE build next to the production library)
_SO = os.environ.get("DETRB_SO") or os.path.join(_HERE, "libdetrb.so")
_SOURCES = ["abi.cu", "igemm.cu", "wgrad.cu", "attention.cu", "elementwise.cu", "matcher.cu", "optim.cu", "gemm_tc.cu", "tma_maps.cu", "wgrad_tc.cu", "pipeline.cu", "attention_tc.cu", "conv_halo.cu", "handle.cu"]
_lib = None
ABI_VERSION = 223          # detrb_version() of the library these ctypes structures / call sites were written for

EXPORTS = [
    "detrb_version", "detrb_last_error", "detrb_check_device", "detrb_set_pdl", "detrb_igemm", "detrb_wgrad", "detrb_attn_fwd",
    "detrb_attn_bwd", "detrb_layernorm_fwd", "detrb_layernorm_bwd", "detrb_add_rowbcast", "detrb_add",
    "detrb_image_to_nhwc4", "detrb_image_to_s2d16", "detrb_f32_to_bf16", "detrb_colsum", "detrb_maxpool_fwd", "detrb_maxpool_bwd",
    "detrb_matcher", "detrb_set_loss", "detrb_adam_clipnorm", "detrb_prep_weight", "detrb_dropout_mask",
    "detrb_set_tc", "detrb_set_tc_conv", "detrb_set_tc_tma_epilogue", "detrb_set_tc_persistent", "detrb_gemm_tc_force", "detrb_prep_weights_multi", "detrb_adam_clipnorm_chunked", "detrb_set_tc_wgrad", "detrb_wgrad_tc_force",
    "detrb_create", "detrb_destroy", "detrb_handle_device", "detrb_handle_set", "detrb_handle_get", "detrb_bind", "detrb_set_wgrad_tile",
    "detrb_attn_dropout_mask", "detrb_normalize_u8", "detrb_image_u8_to_s2d16", "detrb_postprocess", "detrb_accumulate", "detrb_set_tc_attn", "detrb_map_match", "detrb_resize_affine_u8", "detrb_set_tc_stream", "detrb_set_tc_halo", "detrb_set_tc_pair",
]


def sources():
    return [os.path.join(_CSRC, s) for s in _SOURCES if os.path.exists(os.path.join(_CSRC, s))]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -> detr_tensorflow_b200/libdetrb.so (in-tree)."""
    srcs = sources()
    deps = srcs + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    deps.append(os.path.join(os.path.dirname(_HERE), "include", "detrb.h"))
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in deps):
        return _SO
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
           "-t", str(min(8, os.cpu_count() or 1)), "-Xcompiler", "-fPIC", "-shared", "-o", _SO] + os.environ.get("DETRB_NVCC_FLAGS", "").split() + srcs
    if verbose:
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return _SO


class IgemmParams(Structure):
    _fields_ = [
        ("A", c_void_p), ("W", c_void_p), ("M", c_int), ("N", c_int), ("K", c_int), ("lda", c_int), ("ldw", c_int),
        ("batch", c_int), ("IH", c_int), ("IW", c_int), ("Cin", c_int), ("OH", c_int), ("OW", c_int),
        ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad", c_int), ("mode", c_int),
        ("bias", c_void_p), ("residual", c_void_p), ("ldr", c_int), ("mask", c_void_p), ("ldm", c_int),
        ("mask_scale", c_float), ("relu", c_int), ("sigmoid", c_int), ("drop_p", c_float), ("seed", c_uint64),
        ("site", c_uint32), ("seed_ptr", c_void_p), ("C", c_void_p), ("ldc", c_int), ("Cf", c_void_p), ("ldcf", c_int),
        ("out_stride", c_int), ("SH", c_int), ("SW", c_int), ("accumulate", c_int), ("a_kb_rows", c_int),
        ("split", c_int64), ("wsplit", c_int64), ("mask_bits", c_void_p), ("ldmb", c_int), ("out_bits", c_void_p), ("ldob", c_int), ("scratch", c_void_p),
    ]


class WgradParams(Structure):
    _fields_ = [
        ("A", c_void_p), ("lda", c_int), ("dY", c_void_p), ("ldy", c_int), ("M", c_int), ("N", c_int), ("K", c_int),
        ("batch", c_int), ("IH", c_int), ("IW", c_int), ("Cin", c_int), ("OH", c_int), ("OW", c_int),
        ("KH", c_int), ("KW", c_int), ("stride", c_int), ("pad", c_int),
        ("rowscale", c_void_p), ("dW", c_void_p), ("ldw", c_int), ("dbias", c_void_p), ("a_kb_rows", c_int), ("k_mask", c_int),
        ("split", c_int64),
    ]


class PrepDesc(Structure):
    _fields_ = [
        ("master", c_void_p), ("fold", c_void_p), ("Wf", c_void_p), ("Wd", c_void_p),
        ("N", c_int), ("taps", c_int), ("Cin", c_int), ("ldf", c_int), ("ldd", c_int), ("tile_begin", c_int),
    ]


class AttnFwdParams(Structure):
    _fields_ = [
        ("Q", c_void_p), ("K", c_void_p), ("V", c_void_p), ("ldq", c_int), ("ldk", c_int), ("ldv", c_int),
        ("O", c_void_p), ("ldo", c_int), ("lse", c_void_p), ("B", c_int), ("H", c_int), ("Lq", c_int), ("Lk", c_int),
        ("scale", c_float), ("drop_p", c_float), ("seed", c_uint64), ("site", c_uint32), ("seed_ptr", c_void_p),
        ("split", c_int64),
    ]


class AttnBwdParams(Structure):
    _fields_ = [
        ("Q", c_void_p), ("K", c_void_p), ("V", c_void_p), ("O", c_void_p), ("dO", c_void_p),
        ("ldq", c_int), ("ldk", c_int), ("ldv", c_int), ("ldo", c_int), ("lddo", c_int),
        ("lse", c_void_p), ("delta", c_void_p), ("dQ", c_void_p), ("dK", c_void_p), ("dV", c_void_p),
        ("lddq", c_int), ("lddk", c_int), ("lddv", c_int), ("B", c_int), ("H", c_int), ("Lq", c_int), ("Lk", c_int),
        ("scale", c_float), ("drop_p", c_float), ("seed", c_uint64), ("site", c_uint32), ("seed_ptr", c_void_p),
        ("split", c_int64), ("parts", c_int),
    ]


def lib():
    """Load libdetrb.so, (re)building it first when it is missing or older than its sources and nvcc is present (build() is a
    no-op for an up-to-date binary).  A library whose detrb_version() differs from ABI_VERSION is refused: the ctypes structures
    above are passed by reference, so a stale binary would misread their fields silently."""
    global _lib
    if _lib is not None:
        return _lib
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.environ.get("DETRB_SO") and (not os.path.exists(_SO) or (os.path.exists(nvcc) and os.path.isdir(_CSRC))):
        build()
    L = ctypes.CDLL(_SO)
    L.detrb_last_error.restype = c_char_p
    L.detrb_version.restype = c_int
    for name in EXPORTS:
        getattr(L, name)         # raises AttributeError if a declared symbol is missing
    if L.detrb_version() != ABI_VERSION:
        raise RuntimeError("%s reports ABI version %d, this package needs %d: rebuild it (detr_tensorflow_b200._lib.build(force=True))"
                           % (_SO, L.detrb_version(), ABI_VERSION))
    _lib = L
    return L


class DetrbError(RuntimeError):
    pass


def check(rc):
    if rc != 0:
        raise DetrbError("libdetrb error %d: %s" % (rc, lib().detrb_last_error().decode()))
