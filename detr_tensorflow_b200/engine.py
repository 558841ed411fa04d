"""DETR train-step engine: owns device memory (torch tensors) and drives the sm_100a kernels of libdetrb
through the C ABI.  Pure orchestration -- every FLOP of the step runs in csrc/*.cu.

Data layout in HBM
  * activations: NHWC / [tokens, channels] row-major bf16, kept resident for the backward pass
  * parameters : one flat fp32 arena (master weights) + same-shaped arenas for grads / Adam m / Adam v; conv
                 kernels are stored [Cout][tap][Cin] (GEMM K-major), Linear kernels [out][in] as in the reference
  * bf16 weight copies in kernel layouts (forward [N][K], data-gradient [Cin][tap][N]) refreshed after each
                 optimizer step; FrozenBatchNorm2D is folded into them (scale) and into the epilogue bias (shift)
  * precision="parity": every bf16 activation / weight copy is a PAIR of bf16 planes (hi, lo = x - hi; include/detrb.h):
                 all activation buffers are carved from one [2, P] arena so that the lo twin of any view sits exactly P
                 elements after it (one plane stride for every kernel call); the GEMM kernels run three tensor-core passes
                 per product, attention runs in fp32 -- the mode the fp32-tolerance parity tests use

Reference being replaced: detr_tf/networks/{detr,resnet_backbone,transformer,position_embeddings,custom_layers}.py,
detr_tf/loss/{loss,hungarian_matching}.py, detr_tf/optimizers.py, detr_tf/training.py:9-25.
"""
from __future__ import annotations

import math
import os
from collections import OrderedDict

import torch

from . import _lib, ops
from .networks.spec import RESNET_STAGES, head_names, model_params

BF16 = torch.bfloat16
F32 = torch.float32
GROUPS = ("backbone", "transformers", "nlayers")


def _round_up(x, m):
    return (x + m - 1) // m * m


class Slot:
    """A GEMM-able weight (conv or linear) with its optimizer views and bf16 kernel-layout copies."""
    def __init__(self):
        self.name = None
        self.N = self.taps = self.Cin = self.K = 0
        self.master = self.grad = None          # fp32 [N*K] views into the arenas
        self.bias = self.bias_grad = None       # fp32 [N] views (trainable bias) or None
        self.fold = None                        # fp32 [N]: FrozenBN scale folded into the weight
        self.shift = None                       # fp32 [N]: FrozenBN shift used as epilogue bias
        self.Wf = self.Wd = None
        self.ldd = 0
        self.geom = None                        # (kh, kw_real, kw_padded, stride, pad) for convs

    @property
    def epi_bias(self):
        return self.shift if self.shift is not None else self.bias


class Engine:
    def __init__(self, device="cuda", backbone="resnet50", num_classes=92, num_encoder_layers=6,
                 num_decoder_layers=6, num_queries=100, dropout=0.1, seed=0, nb_class=None, precision="bf16"):
        if precision not in ("bf16", "parity"):
            raise ValueError("precision must be 'bf16' (throughput) or 'parity' (bf16 pairs, fp32-class arithmetic)")
        self.precision = precision
        self.paired = precision == "parity"
        self.plane = 0                              # activation plane stride (elements), set by _plan in parity precision
        self.wplane = 0                             # weight-copy plane stride
        self.device = torch.device(device)
        self.lib = _lib.lib()
        if self.device.type != "cuda" and not getattr(_lib, "_EMULATED", False):
            raise RuntimeError("detr_tensorflow_b200 has no CPU path: a B200 (sm_100a) device is required")
        self.handle = None
        if self.device.type == "cuda":
            # one handle per GPU / rank: fails unless the device is compute capability 10.x (there is no fallback path), then makes
            # the device and the handle's kernel-policy switches current for this thread (include/detrb.h, "Handles")
            index = self.device.index if self.device.index is not None else torch.cuda.current_device()
            self.handle = ops.Handle(index)
            self.handle.bind()
            _lib.check(self.lib.detrb_check_device())
        self.backbone_name = backbone
        self.nb_class = nb_class                    # fine-tuning heads (detr.py:94-114): C = nb_class, 'nlayers' group
        self.C = num_classes if nb_class is None else int(nb_class)
        self.nenc, self.ndec, self.Q = num_encoder_layers, num_decoder_layers, num_queries
        self.d, self.H, self.dff = 256, 8, 2048
        self.dropout = dropout
        self.base_seed = seed
        self.spec = model_params(num_classes, backbone, num_encoder_layers, num_decoder_layers, 256, 2048, num_queries,
                                 nb_class=nb_class)
        self._build_params()
        self.plan_key = None
        self._in_backward = False
        self.sites = {}
        self.seed_dev = torch.zeros(1, dtype=torch.int64, device=self.device)
        self.launches = 0
        self.acc = None
        # weight gradients run on a side stream, concurrently with the data-gradient chain (they only feed the optimizer)
        self.overlap_wgrad = self.device.type == "cuda" and not self.paired     # (parity precision: one stream, one arena)
        self.overlap_fwd = os.environ.get("DETRB_NO_FWD_FORK") is None        # forward-pass branches on the side stream
        self.dec_fork = os.environ.get("DETRB_DEC_FORK", "1") != "0"           # decoder: value projection on a second side stream
        self.overlap_dmem = os.environ.get("DETRB_NO_DMEM_FORK") is None      # d(memory) accumulation on the side stream
        self._wstream = None
        self._pending = {}
        self._w_last = None

    # ------------------------------------------------------------------------------------------ parameters
    def _build_params(self):
        dev = self.device
        spec = self.spec
        self.slots = OrderedDict()          # slot name (prefix) -> Slot
        self.vars = []                      # (ref_name, offset, numel, group, kind)
        off = 0

        def alloc(n):
            nonlocal off
            o = off
            off = _round_up(off + n, 64)
            return o
        layout = {}
        for group in GROUPS:
            for name, p in spec.items():
                if p.group != group:
                    continue
                if p.kind == "conv":
                    kh, kw, ci, co = p.shape
                    if name == "backbone/conv1/kernel":
                        n = co * 16 * 16             # space-to-depth form: 4 x 4 taps x 16 channels (see _stem_to_s2d)
                    else:
                        n = co * kh * kw * ci
                else:
                    n = 1
                    for s in p.shape:
                        n *= s
                o = alloc(n)
                layout[name] = (o, n)
                self.vars.append((name, o, n, group, p.kind))
        self.total = off
        self.params = torch.zeros(self.total, dtype=F32, device=dev)
        self.grads = torch.zeros(self.total, dtype=F32, device=dev)
        self.adam_m = torch.zeros(self.total, dtype=F32, device=dev)
        self.adam_v = torch.zeros(self.total, dtype=F32, device=dev)
        self.layout = layout
        T = len(self.vars)
        self.T = T
        self.table = torch.tensor([[o, n] for (_, o, n, _, _) in self.vars], dtype=torch.int64).to(dev)
        self.lr_group = torch.tensor([GROUPS.index(g) for (_, _, _, g, _) in self.vars], dtype=torch.int32).to(dev)
        self.lrs = torch.zeros(8, dtype=F32, device=dev)
        self.group_enabled = torch.zeros(8, dtype=torch.uint8, device=dev)
        self.steps = torch.zeros(8, dtype=torch.int32, device=dev)
        self.norms = torch.zeros(T, dtype=F32, device=dev)
        CH = 8192
        chunks = [[t, o + c, min(CH, n - c)] for t, (_, o, n, _, _) in enumerate(self.vars) for c in range(0, n, CH)]
        self.chunks = torch.tensor(chunks, dtype=torch.int32).to(dev)
        self.nchunks = len(chunks)
        assert self.vars[0][0] == "backbone/conv1/kernel"
        self.stem_chunks = sum(1 for c in chunks if c[0] == 0)      # the stem kernel's chunks come first (optimizer_step)
        self.group_chunks = {}                                       # group -> [lo, hi) of the chunk table (groups are contiguous)
        for gi, g in enumerate(GROUPS):
            idx = [i for i, c in enumerate(chunks) if self.vars[c[0]][3] == g]
            if idx:
                assert idx == list(range(idx[0], idx[-1] + 1))
                self.group_chunks[g] = (idx[0], idx[-1] + 1)
        self.group_range = {}
        for g in GROUPS:
            offs = [(o, n) for (_, o, n, gg, _) in self.vars if gg == g]
            if offs:                                     # 'nlayers' is empty for the include_top=True model
                self.group_range[g] = (offs[0][0], _round_up(offs[-1][0] + offs[-1][1], 64))     # (allocations are 64-aligned)

        def view(arena, name):
            o, n = layout[name]
            return arena[o:o + n]

        # ---- slots
        def conv_slot(prefix, kname, bn_prefix=None, bias_name=None, stride=1, pad=0):
            kh, kw, ci, co = spec[kname].shape
            s = Slot()
            s.name = prefix
            stem = (kname == "backbone/conv1/kernel")
            kwp, cip = (4, 16) if stem else (kw, ci)
            s.N, s.taps, s.Cin = co, (16 if stem else kh * kwp), cip
            s.K = s.taps * s.Cin
            s.master, s.grad = view(self.params, kname), view(self.grads, kname)
            if bias_name:
                s.bias, s.bias_grad = view(self.params, bias_name), view(self.grads, bias_name)
            if bn_prefix:
                s.fold = torch.ones(co, dtype=F32, device=dev)
                s.shift = torch.zeros(co, dtype=F32, device=dev)
                s.bn_prefix = bn_prefix
            s.wf_shape = (co, s.K)
            if not stem:
                s.ldd = _round_up(co, 32)
                s.wd_shape = (ci, s.taps, s.ldd)
            s.geom = (4, 4, 4, 1, 2) if stem else (kh, kw, kwp, stride, pad)
            self.slots[prefix] = s
            return s

        def lin_slot(prefix, kname=None, bname=None):
            kname = kname or prefix + "/kernel"
            bname = bname or prefix + "/bias"
            o_, i_ = spec[kname].shape
            if spec[kname].kind == "dense_w":            # Keras Dense kernel [in, out]; stored [out, in] in the arena
                o_, i_ = i_, o_
            s = Slot()
            s.name = prefix
            s.N, s.taps, s.Cin, s.K = o_, 1, i_, i_
            s.master, s.grad = view(self.params, kname), view(self.grads, kname)
            s.bias, s.bias_grad = view(self.params, bname), view(self.grads, bname)
            s.wf_shape = (o_, i_)
            s.ldd = _round_up(o_, 32)
            s.wd_shape = (i_, 1, s.ldd)
            self.slots[prefix] = s
            return s

        conv_slot("backbone/conv1", "backbone/conv1/kernel", "backbone/bn1", stride=2, pad=3)
        self.blocks = []
        cin = 64
        for li, (nb, d1, d2, stride) in enumerate(RESNET_STAGES[self.backbone_name]):
            for b in range(nb):
                p = f"backbone/layer{li + 1}/{b}"
                st = stride if b == 0 else 1
                blk = dict(prefix=p, cin=cin, d1=d1, d2=d2, stride=st, ds=(b == 0), first=(li == 0 and b == 0))
                blk["c1"] = conv_slot(p + "/conv1", p + "/conv1/kernel", p + "/bn1")
                blk["c2"] = conv_slot(p + "/conv2", p + "/conv2/kernel", p + "/bn2", stride=st, pad=1)
                blk["c3"] = conv_slot(p + "/conv3", p + "/conv3/kernel", p + "/bn3")
                if b == 0:
                    blk["cd"] = conv_slot(p + "/downsample", p + "/downsample_0/kernel", p + "/downsample_1", stride=st)
                self.blocks.append(blk)
                cin = d2
        self.c_feat = cin
        conv_slot("input_proj", "input_proj/kernel", None, "input_proj/bias")

        def mha_slots(p):
            return dict(inp=lin_slot(p + "/in_proj", p + "/in_proj_kernel", p + "/in_proj_bias"),
                        out=lin_slot(p + "/out_proj", p + "/out_proj_kernel", p + "/out_proj_bias"))

        def ln_views(p):
            return dict(g=view(self.params, p + "/gamma"), b=view(self.params, p + "/beta"),
                        dg=view(self.grads, p + "/gamma"), db=view(self.grads, p + "/beta"))
        self.enc = []
        for l in range(self.nenc):
            p = f"transformer/encoder/layer_{l}"
            self.enc.append(dict(sa=mha_slots(p + "/self_attn"), l1=lin_slot(p + "/linear1"), l2=lin_slot(p + "/linear2"),
                                 n1=ln_views(p + "/norm1"), n2=ln_views(p + "/norm2")))
        self.dec = []
        for l in range(self.ndec):
            p = f"transformer/decoder/layer_{l}"
            self.dec.append(dict(sa=mha_slots(p + "/self_attn"), ca=mha_slots(p + "/multihead_attn"),
                                 l1=lin_slot(p + "/linear1"), l2=lin_slot(p + "/linear2"),
                                 n1=ln_views(p + "/norm1"), n2=ln_views(p + "/norm2"), n3=ln_views(p + "/norm3")))
        self.dec_norm = ln_views("transformer/decoder/norm")
        self.h_cls, self.h_b0, self.h_b1, self.h_b2 = (lin_slot(n) for n in head_names(self.nb_class))
        # bf16 kernel-layout weight copies: one tensor per layout, or (parity precision) slices of one [2, PW] arena of pairs
        def numel(shape):
            n = 1
            for v in shape:
                n *= v
            return n
        if self.paired:
            total = sum(_round_up(numel(sh), 64) for s in self.slots.values() for sh in (s.wf_shape, getattr(s, "wd_shape", None)) if sh)
            self.wplane = _round_up(total, 64)
            self._warena = torch.zeros(2 * self.wplane, dtype=BF16, device=dev)
        woff = 0
        for s in self.slots.values():
            for attr, sh in (("Wf", s.wf_shape), ("Wd", getattr(s, "wd_shape", None))):
                if sh is None:
                    continue
                if self.paired:
                    setattr(s, attr, self._warena[woff:woff + numel(sh)].view(*sh))
                    woff += _round_up(numel(sh), 64)
                else:
                    setattr(s, attr, torch.zeros(*sh, dtype=BF16, device=dev))
        self.bn = {n: torch.zeros(p.shape, dtype=F32, device=dev) for n, p in spec.items() if p.kind.startswith("bn_")}
        self.query_embed = torch.zeros(self.Q, self.d, dtype=F32, device=dev)
        self.query_pos = None if self.paired else torch.zeros(self.Q, self.d, dtype=BF16, device=dev)   # parity: lives in the arena (_plan)

    def load_params(self, ref_params):
        """ref_params: {reference name: tensor in the reference layout} (HWIO convs, [out,in] linears)."""
        dev = self.device
        for name, p in self.spec.items():
            t = ref_params[name].detach().to(torch.float32).cpu()
            assert tuple(t.shape) == p.shape, (name, tuple(t.shape), p.shape)
            if p.kind.startswith("bn_"):
                self.bn[name].copy_(t)
            elif p.kind == "embed":
                self.query_embed.copy_(t)
            else:
                o, n = self.layout[name]
                self.params[o:o + n].copy_(self._from_ref_layout(name, t))
        # FrozenBatchNorm2D (custom_layers.py:21-24): scale = w * rsqrt(var + eps); shift = b - mean * scale
        for s in self.slots.values():
            if s.fold is not None:
                pre = s.bn_prefix
                scale = self.bn[pre + "/weight"] * torch.rsqrt(self.bn[pre + "/running_var"] + 1e-5)
                s.fold.copy_(scale)
                s.shift.copy_(self.bn[pre + "/bias"] - self.bn[pre + "/running_mean"] * scale)
        if self.query_pos is not None:
            self._store(self.query_pos, self.query_embed)
        self.refresh_weights()

    # ---- parity precision: bf16 pairs
    def _lo(self, t):
        """the lo-plane twin of an activation view (parity precision only)"""
        return self._arena.as_strided(t.size(), t.stride(), t.storage_offset() + self.plane)

    def _store(self, dst, x):
        """fp32 tensor -> activation storage (a bf16 tensor, or the pair of planes)"""
        x = x.to(self.device, F32)
        hi = x.to(BF16)
        dst.copy_(hi)
        if self.paired:
            self._lo(dst).copy_((x - hi.to(F32)).to(BF16))

    def value(self, t):
        """fp32 value of an activation view (hi + lo in parity precision)"""
        return t.to(F32) + self._lo(t).to(F32) if self.paired else t.to(F32)

    def _from_ref_layout(self, name, t):
        """reference-layout tensor (HWIO conv / [out,in] Linear / [in,out] Dense) -> the flat arena layout of variable `name`"""
        p = self.spec[name]
        t = t.detach().to(torch.float32).cpu()
        if p.kind == "conv":
            t = t.permute(3, 0, 1, 2).contiguous()                      # [co, kh, kw, ci]
            if name == "backbone/conv1/kernel":
                t = self._stem_to_s2d(t)
        elif p.kind == "dense_w":
            t = t.t().contiguous()                                       # [in, out] -> [out, in]
        return t.reshape(-1)

    def export_state(self):
        """everything a resume needs, in the reference's layouts: parameters, Adam moments, per-group iteration counts"""
        out = OrderedDict()
        for name, t in self.export_params().items():
            out["param/" + name] = t
        for (name, _, _, _, _) in self.vars:
            out["adam_m/" + name] = self._to_ref_layout(name, self.adam_m)
            out["adam_v/" + name] = self._to_ref_layout(name, self.adam_v)
        out["adam_steps"] = self.steps[:3].detach().cpu().clone()
        return out

    def load_state(self, state):
        self.load_params(OrderedDict((k[len("param/"):], torch.as_tensor(v)) for k, v in state.items() if k.startswith("param/")))
        for (name, o, n, _, _) in self.vars:
            for arena, pre in ((self.adam_m, "adam_m/"), (self.adam_v, "adam_v/")):
                if pre + name in state:
                    arena[o:o + n].copy_(self._from_ref_layout(name, torch.as_tensor(state[pre + name])))
        if "adam_steps" in state:
            self.steps[:3] = torch.as_tensor(state["adam_steps"]).to(self.steps.dtype)

    def export_params(self):
        out = OrderedDict()
        for name, p in self.spec.items():
            if p.kind.startswith("bn_"):
                out[name] = self.bn[name].detach().cpu().clone()
            elif p.kind == "embed":
                out[name] = self.query_embed.detach().cpu().clone()
            else:
                out[name] = self._to_ref_layout(name, self.params)
        return out

    def _to_ref_layout(self, name, arena):
        p = self.spec[name]
        o, n = self.layout[name]
        t = arena[o:o + n].detach().cpu().clone()
        if p.kind == "conv":
            kh, kw, ci, co = p.shape
            if name == "backbone/conv1/kernel":
                t = self._stem_from_s2d(t.reshape(co, 4, 4, 16))
            else:
                t = t.reshape(co, kh, kw, ci)
            return t.permute(1, 2, 3, 0).contiguous()
        if p.kind == "dense_w":
            return t.reshape(p.shape[1], p.shape[0]).t().contiguous()
        return t.reshape(p.shape)

    @staticmethod
    def _stem_index():
        """(ta, tb, ch) of the 4x4x16 space-to-depth kernel <-> (kh, kw, c) of the 7x7x3 kernel (resnet_backbone.py:11):
        input row 2(oy + ta - 2) + ry = 2 oy - 3 + kh  =>  kh = 2 ta + ry - 1, likewise kw; ch = (ry*2+rx)*3 + c."""
        idx = []
        for ta in range(4):
            for tb in range(4):
                for ry in range(2):
                    for rx in range(2):
                        kh, kw = 2 * ta + ry - 1, 2 * tb + rx - 1
                        if 0 <= kh <= 6 and 0 <= kw <= 6:
                            for c in range(3):
                                idx.append((ta, tb, (ry * 2 + rx) * 3 + c, kh, kw, c))
        return idx

    def _stem_to_s2d(self, t):                      # t [co, 7, 7, 3] -> [co, 4, 4, 16]
        out = torch.zeros(t.shape[0], 4, 4, 16, dtype=t.dtype)
        for ta, tb, ch, kh, kw, c in Engine._stem_index():
            out[:, ta, tb, ch] = t[:, kh, kw, c]
        return out

    def _stem_from_s2d(self, t):                    # t [co, 4, 4, 16] -> [co, 7, 7, 3]
        out = torch.zeros(t.shape[0], 7, 7, 3, dtype=t.dtype)
        for ta, tb, ch, kh, kw, c in Engine._stem_index():
            out[:, kh, kw, c] = t[:, ta, tb, ch]
        return out

    def export_grads(self):
        return OrderedDict((name, self._to_ref_layout(name, self.grads)) for (name, _, _, _, _) in self.vars)

    def refresh_weights(self):
        """fp32 master -> bf16 kernel-layout copies (FrozenBN scale folded in): ONE multi-tensor launch."""
        if getattr(self, "_prep_table", None) is None:
            import ctypes
            import numpy as np
            n = len(self.slots)
            arr = (_lib.PrepDesc * n)()
            begin = 0
            for i, s in enumerate(self.slots.values()):
                d = arr[i]
                d.master, d.fold = s.master.data_ptr(), (s.fold.data_ptr() if s.fold is not None else None)
                d.Wf, d.Wd = s.Wf.data_ptr(), (s.Wd.data_ptr() if s.Wd is not None else None)
                d.N, d.taps, d.Cin, d.ldf, d.ldd, d.tile_begin = s.N, s.taps, s.Cin, s.K, s.ldd, begin
                assert s.Cin % 2 == 0 and s.K % 2 == 0 and (s.Wd is None or s.ldd % 2 == 0)
                begin += s.taps * ((s.N + 63) // 64) * ((s.Cin + 63) // 64)          # 64x64 tiles (optim.cu)
            raw = np.frombuffer(ctypes.string_at(ctypes.addressof(arr), ctypes.sizeof(arr)), dtype=np.uint8).copy()
            self._prep_table = torch.from_numpy(raw).to(self.device)
            self._prep_n, self._prep_tiles = n, begin
        if hasattr(self.lib, "detrb_prep_weights_multi"):
            ops.prep_weights_multi(self._prep_table, self._prep_n, self._prep_tiles, wsplit=self.wplane)
            self.launches += 1
        else:
            assert not self.paired
            for s in self.slots.values():
                ops.prep_weight(s.master, s.fold, s.N, s.taps, s.Cin, s.Wf, s.K, s.Wd, s.ldd)
                self.launches += 1

    # ------------------------------------------------------------------------------------------ plan / buffers
    def _plan(self, B, H, W):
        key = (B, H, W)
        if self.plan_key == key:
            return
        dev = self.device
        self.plan_key = key
        self.B = B

        def o(n, k, s, p):
            return (n + 2 * p - k) // s + 1
        a = {}
        pending = []                                # parity precision: bf16 buffers are carved from one arena of pairs afterwards

        def buf(name, *shape, dtype=BF16, zero=False):
            if self.paired and dtype == BF16:
                pending.append((name, shape))
                return None
            a[name] = (torch.zeros if zero else torch.empty)(*shape, dtype=dtype, device=dev)
            return a[name]
        self.a = a
        self.H0, self.W0 = H, W
        h1, w1 = o(H, 7, 2, 3), o(W, 7, 2, 3)
        h2, w2 = o(h1, 3, 2, 1), o(w1, 3, 2, 1)
        self.hw_stem, self.hw_pool = (h1, w1), (h2, w2)
        self.hw_s2d = ((H + 1) // 2, (W + 1) // 2)
        # stem as a sliding-window GEMM: the space-to-depth image carries the stem's zero padding explicitly (2 pixels above / left,
        # 1 below / right), so the 4 taps (tb) of window row ta at output pixel q are the 64 contiguous elements starting at flat
        # pixel q + ta*WP -- the stem output (and its gradient) live at the same [HP, WP] pitch; the 3 extra rows / columns per
        # image hold wrapped-window values that nothing reads (maxpool_fwd) and zeros in the gradient (maxpool_bwd)
        self.hw_pad = (self.hw_s2d[0] + 3, self.hw_s2d[1] + 3)
        HP, WP = self.hw_pad
        self.M_stem = B * HP * WP
        # + the reach of the last window (3 rows + 4 pixels): zero, never written
        buf("s2d", self.M_stem + 3 * WP + 8, 16, zero=True)
        buf("stem", B, HP, WP, 64)
        buf("pool", B, h2, w2, 64)
        buf("pool_arg", B, h2, w2, 64, dtype=torch.uint8)
        hh, ww = h2, w2
        max_elems = self.M_stem * 64
        for i, blk in enumerate(self.blocks):
            st = blk["stride"]
            ho, wo = (o(hh, 3, st, 1), o(ww, 3, st, 1)) if st > 1 else (hh, ww)
            blk["in_hw"], blk["out_hw"] = (hh, ww), (ho, wo)
            buf(f"b{i}_a1", B, hh, ww, blk["d1"])
            buf(f"b{i}_a2", B, ho, wo, blk["d1"])
            if blk["ds"]:
                buf(f"b{i}_idn", B, ho, wo, blk["d2"])
            buf(f"b{i}_out", B, ho, wo, blk["d2"])
            # 1-bit ReLU masks of the three activations (written by the forward epilogues, read by the data gradients instead of
            # the bf16 activations: 1/16 of the bytes)
            buf(f"b{i}_a1_bits", B * hh * ww, blk["d1"] // 8, dtype=torch.uint8)
            buf(f"b{i}_a2_bits", B * ho * wo, blk["d1"] // 8, dtype=torch.uint8)
            buf(f"b{i}_out_bits", B * ho * wo, blk["d2"] // 8, dtype=torch.uint8)
            max_elems = max(max_elems, B * hh * ww * max(blk["cin"], blk["d1"]), B * ho * wo * blk["d2"])
            hh, ww = ho, wo
        self.fh, self.fw = hh, ww
        S = hh * ww
        self.S = S
        M, Mq, d, dff, Q = B * S, B * self.Q, self.d, self.dff, self.Q
        self.M, self.Mq = M, Mq
        # position embedding: input independent (all-False mask, detr.py:172) -> computed once, fp32 then bf16
        buf("pos", S, d)
        if self.paired:
            buf("query_pos", self.Q, d)
        buf("src", M, d)
        buf("srcp", M, d)
        for l in range(self.nenc):
            for nm, shp in (("qk", (M, 2 * d)), ("v", (M, d)), ("o", (M, d)), ("pre1", (M, d)), ("y1", (M, d)),
                            ("h", (M, dff)), ("pre2", (M, d)), ("y2", (M, d)), ("y2p", (M, d))):
                buf(f"e{l}_{nm}", *shp)
            buf(f"e{l}_lse", B * self.H * S, dtype=F32)
            for nm in ("mean1", "rstd1", "mean2", "rstd2"):
                buf(f"e{l}_{nm}", M, dtype=F32)
        buf("tgt0", Mq, d, zero=True)
        for l in range(self.ndec):
            for nm, shp in (("tq", (Mq, d)), ("qk", (Mq, 2 * d)), ("v", (Mq, d)), ("o", (Mq, d)), ("pre1", (Mq, d)),
                            ("t1", (Mq, d)), ("t1q", (Mq, d)), ("q2", (Mq, d)), ("k2", (M, d)), ("v2", (M, d)),
                            ("o2", (Mq, d)), ("pre2", (Mq, d)), ("t2", (Mq, d)), ("h", (Mq, dff)), ("pre3", (Mq, d))):
                buf(f"d{l}_{nm}", *shp)
            buf(f"d{l}_lse1", B * self.H * Q, dtype=F32)
            buf(f"d{l}_lse2", B * self.H * Q, dtype=F32)
            for nm in ("mean1", "rstd1", "mean2", "rstd2", "mean3", "rstd3"):
                buf(f"d{l}_{nm}", Mq, dtype=F32)
        L = self.ndec
        # the layer outputs t3 are rows of ONE buffer: the shared final norm (transformer.py:122-126) of all layers is one launch
        # after the stack, and its backward one launch before it (d{l}_t3 below are row slices)
        buf("t3_all", L * Mq, d)
        buf("g_t3f", L * Mq, d)
        buf("meanf", L * Mq, dtype=F32)
        buf("rstdf", L * Mq, dtype=F32)
        buf("hs", L * Mq, d)
        buf("logits", L * Mq, self.C, dtype=F32)
        buf("hb1", L * Mq, d)
        buf("hb2", L * Mq, d)
        buf("boxes", L * Mq, 4, dtype=F32)
        # matcher / loss
        P = L * B
        buf("p_indices", P, Q, dtype=torch.int64)
        buf("t_indices", P, Q, dtype=torch.int64)
        buf("p_selector", P, Q, dtype=torch.uint8)
        buf("match", P, Q, dtype=torch.int32)
        buf("status", P, dtype=torch.int32)
        buf("loss_sums", L, 8, dtype=F32)
        buf("losses", L, 6, dtype=F32)
        buf("total", 1, dtype=F32)
        buf("t_bbox", B, 100, 4, dtype=F32)
        buf("t_class", B, 100, 1, dtype=torch.int64)
        buf("images", B, H, W, 3, dtype=F32)
        self.u8_input = False                   # True: a["images_u8"] holds raw uint8 frames, normalised inside the stem's input stage
        self.ld_dl = _round_up(self.C, 32)
        buf("d_logits", L * Mq, self.ld_dl)
        buf("d_boxpre", L * Mq, 32)
        # backward scratch
        buf("g_hs", L * Mq, d)
        buf("g_hb1", L * Mq, d)
        buf("g_hb2", L * Mq, d)
        for nm, shp in (("gq_a", (Mq, d)), ("gq_b", (Mq, d)), ("gq_c", (Mq, d)), ("gq_d", (Mq, d)), ("gq_qkv", (Mq, 3 * d)),
                        ("gq_n0", (Mq, d)), ("gq_n1", (Mq, d)), ("gm_n0", (M, d)), ("gm_n1", (M, d)),
                        ("gq_h", (Mq, dff)), ("gq_q2", (Mq, d)), ("gm_k2", (M, d)), ("gm_v2", (M, d)),
                        ("g_mem", (M, d)), ("gm_a", (M, d)), ("gm_b", (M, d)), ("gm_c", (M, d)), ("gm_d", (M, d)),
                        ("gm_qkv", (M, 3 * d)), ("gm_h", (M, dff))):
            buf(nm, *shp)
        buf("delta", B * self.H * max(S, Q), dtype=F32)
        buf("delta2", L * B * self.H * Q, dtype=F32)             # cross attention: per layer (its dK/dV kernel runs on the side stream)
        buf("gq_d2", L * Mq, d)
        buf("g_x", max_elems)
        buf("g_y", max_elems)
        buf("g_1", max_elems)
        buf("g_2", max_elems)
        if not self.paired:                      # dense workspace of the stride-2 scatter forms (detrb_igemm_t.scratch)
            buf("g_s", max_elems + 64)
        if self.paired:
            def numel(shape):
                n = 1
                for v in shape:
                    n *= int(v)
                return n
            self.plane = _round_up(sum(_round_up(numel(sh), 128) for _, sh in pending), 128)
            self._arena = torch.zeros(2 * self.plane, dtype=BF16, device=dev)         # zero: s2d padding, tgt0, both planes
            off = 0
            for name, sh in pending:
                a[name] = self._arena[off:off + numel(sh)].view(*[int(v) for v in sh])
                off += _round_up(numel(sh), 128)
            self.query_pos = a["query_pos"]
            self._store(self.query_pos, self.query_embed)
        self.pos = a["pos"]
        for l in range(self.ndec):
            a[f"d{l}_t3"] = a["t3_all"][l * Mq:(l + 1) * Mq]
        self._store(self.pos, self._pos_embedding(hh, ww))
        self.normalisers = None

    def _pos_embedding(self, h, w):
        """position_embeddings.py:23-50 with an all-False mask -> [h*w, 256] fp32 (host, once per shape)."""
        eps, scale, npf, T = 1e-6, 2 * math.pi, 128, 10000.0
        y = torch.arange(1, h + 1, dtype=F32).view(h, 1).expand(h, w)
        x = torch.arange(1, w + 1, dtype=F32).view(1, w).expand(h, w)
        y = y / (y[-1:, :] + eps) * scale
        x = x / (x[:, -1:] + eps) * scale
        dim_t = torch.arange(npf, dtype=F32)
        dim_t = T ** (2 * torch.div(dim_t, 2, rounding_mode="floor") / npf)
        px, py = x[..., None] / dim_t, y[..., None] / dim_t
        px = torch.stack([px[..., 0::2].sin(), px[..., 1::2].cos()], dim=3).reshape(h, w, -1)
        py = torch.stack([py[..., 0::2].sin(), py[..., 1::2].cos()], dim=3).reshape(h, w, -1)
        return torch.cat([py, px], dim=2).reshape(h * w, 2 * npf)

    # ------------------------------------------------------------------------------------------ op helpers
    def _mark(self, name):
        """section boundary for profile_step(): a CUDA event on the launch stream (no-op otherwise)"""
        marks = getattr(self, "_marks", None)
        if marks is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            marks.append((name, ev))

    def profile_step(self, background_class, clipnorm, repeats=3):
        """eager train steps with CUDA events at section boundaries -> {section: ms} (median over repeats)"""
        import statistics
        acc = {}
        for _ in range(repeats):
            self._marks = []
            self._mark("start")
            self.train_step(background_class, clipnorm)
            torch.cuda.synchronize()
            marks, self._marks = self._marks, None
            for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
                acc.setdefault(n1, []).append(e0.elapsed_time(e1))
        return {k: statistics.median(v) for k, v in acc.items()}

    def _site(self, name):
        if name not in self.sites:
            self.sites[name] = len(self.sites) + 1
        return self.sites[name]

    def _drop(self, name):
        if self.training and self.dropout > 0:
            return dict(drop_p=self.dropout, seed=self.base_seed, site=self._site(name), seed_ptr=self.seed_dev)
        return {}

    # ---- side-stream weight gradients with write-after-read tracking on the scratch buffers they read
    @staticmethod
    def _key(t):
        return t.untyped_storage().data_ptr()

    def _on_wstream(self, fn, reads, low=False):
        if not (self.overlap_wgrad and self._in_backward):
            return fn()
        if self._wstream is None:
            self._wstream = self._side_stream()
        side = self._wstream
        if low and self._low_stream() is not None:
            side = self._lstream
        main = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(main)
        side.wait_event(ev)                          # operands produced by everything enqueued so far
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        for t in reads:
            self._pending[self._key(t)] = done
        if side is self._wstream:
            self._w_last = done
        else:
            self._l_last = done

    # Stream priorities (kernel nodes of a captured graph keep the priority of the stream they were captured on).
    # DETRB_STREAM_PRIO: 0 (default) = none, 1 = the main chain's thread blocks are placed first, 2 = the side streams' are.
    # Measured on the full-size step (tests/time_step_env.py): 1 costs 0.4 ms -- starved weight gradients pile up behind the
    # backward pass, where they run alone at low occupancy.
    # DETRB_LOWPRIO (default 1): a third, LOWER-priority stream for the bulk work that shadows the decoder -- all layers'
    # cross-attention K/V projections of the memory in the forward pass; the cross-attention dK/dV kernel, its weight gradients and
    # the d(memory) accumulation in the backward pass -- 8400-row kernels that otherwise take the SM slots the decoder's 7..64-CTA
    # kernels are waiting for (the decoder is a latency chain; this work is only needed layers later).
    @staticmethod
    def _prio_mode():
        return int(os.environ.get("DETRB_STREAM_PRIO", "0"))

    @staticmethod
    def _lowprio():
        return os.environ.get("DETRB_LOWPRIO", "1") != "0"

    def _side_stream(self):
        return torch.cuda.Stream(priority=-1 if (self._prio_mode() == 2 or self._lowprio()) else 0)

    def _chain_stream(self):
        return torch.cuda.Stream(priority=-1 if (self._prio_mode() == 1 or self._lowprio()) else 0)

    def _low_stream(self):
        if getattr(self, "_lstream", None) is None:
            self._lstream = torch.cuda.Stream(priority=0) if self._lowprio() else None
        return self._lstream

    def _fork(self, fn, lane=0):
        """Run fn() on a side stream, ordered after everything enqueued on the main stream so far; returns the handle for
        _join().  For branches of the forward pass whose inputs are ready and whose outputs are only needed later.
        lane 1 is a second side stream for short branches that must not queue behind a long one on lane 0."""
        if not (self.overlap_wgrad and self.overlap_fwd):
            fn()
            return None
        if self._wstream is None:
            self._wstream = self._side_stream()
        side = self._wstream
        if lane == 1:
            if getattr(self, "_wstream2", None) is None:
                self._wstream2 = self._side_stream()
            side = self._wstream2
        elif lane == 2 and self._low_stream() is not None:
            side = self._lstream
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        side.wait_event(ev)
        with torch.cuda.stream(side):
            fn()
            done = torch.cuda.Event()
            done.record(side)
        return done

    @staticmethod
    def _join(done):
        if done is not None:
            torch.cuda.current_stream().wait_event(done)

    def _before_write(self, *tensors):
        """a main-stream kernel is about to overwrite these buffers: wait for side-stream readers still pending on them"""
        if not self._pending:
            return
        for t in tensors:
            if t is None:
                continue
            ev = self._pending.pop(self._key(t), None)
            if ev is not None:
                torch.cuda.current_stream().wait_event(ev)

    def _join_wgrad(self):
        if self._w_last is not None:
            torch.cuda.current_stream().wait_event(self._w_last)
        self._w_last = None
        if getattr(self, "_l_last", None) is not None:
            torch.cuda.current_stream().wait_event(self._l_last)
        self._l_last = None
        self._pending.clear()

    def _lin(self, A, W, M, N, K, ldw, out=None, ldc=None, lda=None, **kw):
        """plain GEMM  out[M,N] = A[M,K] . W[N,K]^T (+epilogue)"""
        self.launches += 1
        self._before_write(out, kw.get("Cf"))
        ops.igemm(A, W, M, N, K, lda or K, ldw, ops.plain_geom(M, K), C=out, ldc=(ldc if ldc is not None else N),
                  split=self.plane, wsplit=self.wplane, **kw)

    def _conv_geom(self, s, B, ih, iw, oh, ow, mode=0):
        kh, kw, kwp, stride, pad = s.geom
        return dict(batch=B, IH=ih, IW=iw, Cin=s.Cin, OH=oh, OW=ow, KH=kh, KW=kwp, stride=stride, pad=pad, mode=mode)

    def _conv_fwd(self, s, x, ihw, ohw, out, relu=True, residual=None, out_bits=None):
        B = self.B
        g = self._conv_geom(s, B, ihw[0], ihw[1], ohw[0], ohw[1])
        self.launches += 1
        self._probed(s.name, lambda: ops.igemm(x, s.Wf, B * ohw[0] * ohw[1], s.N, s.K, s.Cin, s.K, g, bias=s.epi_bias, residual=residual,
                                               ldr=s.N, relu=relu, C=out, ldc=s.N, split=self.plane, wsplit=self.wplane,
                                               out_bits=out_bits, ldob=s.N // 8))

    def _probed(self, name, fn):
        """bench.py: CUDA-event timing of single launches (events on the stream the launch goes to); plain call otherwise"""
        if name not in (getattr(self, "probe_names", None) or ()):
            return fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        self.probe_events.setdefault(name, []).append((e0, e1))

    def _conv_dgrad(self, s, dy, ihw, ohw, out, mask_bits=None, residual=None):
        """data gradient of conv `s` (input ihw -> output ohw): out[B,ih,iw,Cin] from dy[B,oh,ow,N]; mask_bits: the 1-bit ReLU mask
        of the conv's input activation."""
        B = self.B
        kh, kw, kwp, stride, pad = s.geom
        M = B * ihw[0] * ihw[1]
        g = dict(batch=B, IH=ohw[0], IW=ohw[1], Cin=s.ldd, OH=ihw[0], OW=ihw[1], KH=kh, KW=kwp, stride=stride, pad=pad, mode=1)
        self.launches += 1
        self._before_write(out)
        self._probed(s.name + "#dgrad", lambda: ops.igemm(dy, s.Wd, M, s.Cin, s.taps * s.ldd, s.N, s.taps * s.ldd, g, mask_bits=mask_bits,
                                                          ldmb=s.Cin // 8, mask_scale=1.0, residual=residual, ldr=s.Cin, C=out, ldc=s.Cin,
                                                          split=self.plane, wsplit=self.wplane,
                                                          scratch=self.a.get("g_s") if stride > 1 else None))

    def _conv_wgrad(self, s, x, dy, ihw, ohw):
        B = self.B
        g = self._conv_geom(s, B, ihw[0], ihw[1], ohw[0], ohw[1])
        self.launches += 1
        self._on_wstream(lambda: self._probed(s.name + "#wgrad", lambda: ops.wgrad(
            x, s.Cin, dy, s.N, B * ohw[0] * ohw[1], s.N, s.K, g, s.grad, s.K, rowscale=s.fold, dbias=s.bias_grad, split=self.plane)), (x, dy))

    def _lin_wgrad(self, s, x, dy, M, ldy=None, n_off=0, n_rows=None, lda=None, low=False):
        """dW[n_off:n_off+n_rows] += dy^T x ; dbias likewise"""
        n_rows = n_rows or s.N
        self.launches += 1
        self._on_wstream(lambda: ops.wgrad(x, lda or s.K, dy, ldy or n_rows, M, n_rows, s.K, ops.plain_geom(M, s.K),
                                           s.grad[n_off * s.K:], s.K, dbias=s.bias_grad[n_off:], split=self.plane), (x, dy), low=low)

    def _ln_fwd(self, x, n, y, mean, rstd, M, y2=None, pos=None, S=1):
        self.launches += 1
        ops.layernorm_fwd(x, n["g"], n["b"], y, y2, pos, S, mean, rstd, M, split=self.plane)

    def _ln_bwd(self, dy, dy2, x, n, mean, rstd, dx, dx_drop, M, drop_name=None):
        self.launches += 1
        self._before_write(dx, dx_drop)
        d = self._drop(drop_name) if drop_name else {}
        ops.layernorm_bwd(dy, dy2, x, n["g"], mean, rstd, dx, dx_drop, d.get("drop_p", 0.0), d.get("seed", 0),
                          d.get("site", 0), d.get("seed_ptr"), n["dg"], n["db"], M, split=self.plane)

    # ------------------------------------------------------------------------------------------ forward
    def forward(self, images, training=False):
        """images: [B,H,W,3] float32 (NHWC like the reference, training.py:18).  Returns the reference's output
        dict {'pred_logits','pred_boxes','aux'} as fp32 device tensors (detr.py:190-204)."""
        B, H, W, _ = images.shape
        self._plan(B, H, W)
        self.training = bool(training)
        self._stage_images(images)
        self._ensure_weights()
        self._forward_impl()
        return self.outputs()

    def set_input_normalisation(self, normalized_method):
        """uint8 input frames are normalised on device with the reference's normalized_images arithmetic (table lookup)"""
        from .data.processing import device_lut
        self.input_lut = device_lut(normalized_method, self.device)
        self.input_method = normalized_method

    def _stage_images(self, images, direct=False):
        """images -> the engine's resident input.  direct (training.fit / run_train_step: `images` is a contiguous tensor on this
        device): the space-to-depth conversion that opens the forward pass runs NOW, reading `images` where it lies, and the replayed
        launch sequence starts behind it (s2d_staged: _forward_impl skips the conversion) -- instead of a 100 MB device-to-device copy
        into the resident image buffer in front of every step (env DETRB_STAGE_S2D=0: the copy)."""
        a = self.a
        self.s2d_staged = False
        if (direct and images.device == self.device and images.is_contiguous() and tuple(images.shape) == tuple(a["images"].shape)
                and images.dtype in (torch.uint8, F32) and os.environ.get("DETRB_STAGE_S2D", "1") != "0"):
            B = self.B
            HP, WP = self.hw_pad
            if images.dtype == torch.uint8:
                if getattr(self, "input_lut", None) is None:
                    self.set_input_normalisation("torch_resnet")
                ops.image_u8_to_s2d16(images, self.input_lut[0], self.input_lut[1], a["s2d"], B, self.H0, self.W0, 2, 2, HP, WP,
                                      split=self.plane)
            else:
                ops.image_to_s2d16(images, a["s2d"], B, self.H0, self.W0, 2, 2, HP, WP, split=self.plane)
            self.launches += 1
            self.u8_input = images.dtype == torch.uint8
            self.s2d_staged = True
            return
        if images.dtype == torch.uint8:
            if getattr(self, "input_lut", None) is None:
                self.set_input_normalisation("torch_resnet")
            if "images_u8" not in a:
                a["images_u8"] = torch.empty(a["images"].shape, dtype=torch.uint8, device=self.device)
            if images.data_ptr() != a["images_u8"].data_ptr():
                a["images_u8"].copy_(images, non_blocking=True)
            self.u8_input = True
        else:
            if images.device != self.device or images.data_ptr() != a["images"].data_ptr():
                a["images"].copy_(images, non_blocking=True)
            self.u8_input = False

    def outputs(self):
        a, L, B, Q = self.a, self.ndec, self.B, self.Q
        lg = a["logits"].view(L, B, Q, self.C)
        bx = a["boxes"].view(L, B, Q, 4)
        return {"pred_logits": lg[L - 1], "pred_boxes": bx[L - 1],
                "aux": [{"pred_logits": lg[i], "pred_boxes": bx[i]} for i in range(L - 1)]}

    def _forward_impl(self):
        a, B = self.a, self.B
        d, dff, S, Q, M, Mq, Hh = self.d, self.dff, self.S, self.Q, self.M, self.Mq, self.H
        scale = float(d // Hh) ** -0.5
        # ---------------- backbone (resnet_backbone.py:20-32)
        # stem: space-to-depth(2) turns the 7x7/s2 conv into a dense 4x4/s1 conv over 16-channel pixels = a sliding-window GEMM
        HP, WP = self.hw_pad
        if getattr(self, "s2d_staged", False):
            pass                              # a["s2d"] was filled from the caller's batch by _stage_images(direct=True)
        elif self.u8_input:                   # data/processing.py:6-23 fused into the layout change (no fp32 image in HBM)
            ops.image_u8_to_s2d16(a["images_u8"], self.input_lut[0], self.input_lut[1], a["s2d"], B, self.H0, self.W0, 2, 2, HP, WP,
                                  split=self.plane)
            self.launches += 1
        else:
            ops.image_to_s2d16(a["images"], a["s2d"], B, self.H0, self.W0, 2, 2, HP, WP, split=self.plane)
            self.launches += 1
        stem = self.slots["backbone/conv1"]
        # one plain GEMM [B*HP*WP, 256] x [64, 256]^T whose A rows are overlapping 128-byte windows of the padded image
        self.launches += 1
        ops.igemm(a["s2d"], stem.Wf, self.M_stem, stem.N, stem.K, 16, stem.K, ops.plain_geom(self.M_stem, stem.K),
                  bias=stem.epi_bias, relu=True, C=a["stem"], ldc=stem.N, a_kb_rows=WP, split=self.plane, wsplit=self.wplane)
        ops.maxpool_fwd(a["stem"], a["pool"], a["pool_arg"], B, self.hw_stem[0], self.hw_stem[1], 64, self.hw_pool[0], self.hw_pool[1],
                        XH=HP, XW=WP, split=self.plane)
        self.launches += 1
        x = a["pool"]
        for i, blk in enumerate(self.blocks):
            ihw, ohw = blk["in_hw"], blk["out_hw"]
            ds = None
            if blk["ds"]:                                     # the shortcut convolution runs beside conv1 -> conv2
                ds = self._fork(lambda: self._conv_fwd(blk["cd"], x, ihw, ohw, a[f"b{i}_idn"], relu=False))
                idn = a[f"b{i}_idn"]
            else:
                idn = x
            self._conv_fwd(blk["c1"], x, ihw, ihw, a[f"b{i}_a1"], out_bits=a[f"b{i}_a1_bits"])
            self._conv_fwd(blk["c2"], a[f"b{i}_a1"], ihw, ohw, a[f"b{i}_a2"], out_bits=a[f"b{i}_a2_bits"])
            self._join(ds)
            self._conv_fwd(blk["c3"], a[f"b{i}_a2"], ohw, ohw, a[f"b{i}_out"], relu=True, residual=idn, out_bits=a[f"b{i}_out_bits"])
            x = a[f"b{i}_out"]
        self.feat = x
        self._mark("fwd_backbone")
        # ---------------- input_proj (detr.py:44,175) + pos add
        ip = self.slots["input_proj"]
        self._lin(x, ip.Wf, M, d, ip.K, ip.K, out=a["src"], bias=ip.bias)
        ops.add_rowbcast(a["src"], self.pos, a["srcp"], M, S, d, split=self.plane)
        self.launches += 1
        # ---------------- encoder (transformer.py:157-179)
        xin, xinp = a["src"], a["srcp"]
        for l, E in enumerate(self.enc):
            e = lambda n: a[f"e{l}_{n}"]
            W = E["sa"]["inp"]
            vj = self._fork(lambda: self._lin(xin, W.Wf[2 * d:], M, d, d, d, out=e("v"), bias=W.bias[2 * d:]))
            self._lin(xinp, W.Wf, M, 2 * d, d, d, out=e("qk"), bias=W.bias)
            self._join(vj)
            self.launches += 1
            self._probed(f"e{l}_attn#fwd", lambda: ops.attn_fwd(e("qk"), e("qk")[:, d:], e("v"), 2 * d, 2 * d, d, e("o"), d, e("lse"),
                                                                 B, Hh, S, S, scale, **self._attn_drop(f"e{l}_attn")))
            Wo = E["sa"]["out"]
            self._lin(e("o"), Wo.Wf, M, d, d, d, out=e("pre1"), bias=Wo.bias, residual=xin, ldr=d, **self._drop(f"e{l}_do1"))
            self._ln_fwd(e("pre1"), E["n1"], e("y1"), e("mean1"), e("rstd1"), M)
            self._lin(e("y1"), E["l1"].Wf, M, dff, d, d, out=e("h"), bias=E["l1"].bias, relu=True, **self._drop(f"e{l}_dh"))
            self._lin(e("h"), E["l2"].Wf, M, d, dff, dff, out=e("pre2"), bias=E["l2"].bias, residual=e("y1"), ldr=d,
                      **self._drop(f"e{l}_do2"))
            self._ln_fwd(e("pre2"), E["n2"], e("y2"), e("mean2"), e("rstd2"), M, y2=e("y2p"), pos=self.pos, S=S)
            xin, xinp = e("y2"), e("y2p")
        mem, memp = xin, xinp
        self.mem, self.memp = mem, memp
        self._mark("fwd_encoder")
        # ---------------- decoder (transformer.py:207-234, 104-133)
        tgt = a["tgt0"]
        # the cross-attention key / value projections of the encoder memory do not depend on the decoder state: all layers'
        # run on the side stream while the decoder's own chain proceeds
        kv_ready = []
        for l, D in enumerate(self.dec):
            Wc = D["ca"]["inp"]

            def kv(l=l, Wc=Wc):
                self._lin(memp, Wc.Wf[d:], M, d, d, d, out=a[f"d{l}_k2"], bias=Wc.bias[d:])
                self._lin(mem, Wc.Wf[2 * d:], M, d, d, d, out=a[f"d{l}_v2"], bias=Wc.bias[2 * d:])
            kv_ready.append(self._fork(kv, lane=2))
        for l, D in enumerate(self.dec):
            t = lambda n: a[f"d{l}_{n}"]
            W = D["sa"]["inp"]
            vj = None
            if self.dec_fork:                             # value projection beside the query/key projection (as in the encoder)
                vj = self._fork(lambda: self._lin(tgt, W.Wf[2 * d:], Mq, d, d, d, out=t("v"), bias=W.bias[2 * d:]), lane=1)
            if l == 0:                                    # (later layers: tq is the second output of the previous layer's LN3)
                ops.add_rowbcast(tgt, self.query_pos, t("tq"), Mq, Q, d, split=self.plane)
                self.launches += 1
            self._lin(t("tq"), W.Wf, Mq, 2 * d, d, d, out=t("qk"), bias=W.bias)
            if self.dec_fork:
                self._join(vj)
            else:
                self._lin(tgt, W.Wf[2 * d:], Mq, d, d, d, out=t("v"), bias=W.bias[2 * d:])
            self.launches += 1
            ops.attn_fwd(t("qk"), t("qk")[:, d:], t("v"), 2 * d, 2 * d, d, t("o"), d, t("lse1"), B, Hh, Q, Q, scale,
                         **self._attn_drop(f"d{l}_attn1"))
            Wo = D["sa"]["out"]
            self._lin(t("o"), Wo.Wf, Mq, d, d, d, out=t("pre1"), bias=Wo.bias, residual=tgt, ldr=d, **self._drop(f"d{l}_do1"))
            self._ln_fwd(t("pre1"), D["n1"], t("t1"), t("mean1"), t("rstd1"), Mq, y2=t("t1q"), pos=self.query_pos, S=Q)
            W = D["ca"]["inp"]
            self._lin(t("t1q"), W.Wf, Mq, d, d, d, out=t("q2"), bias=W.bias)
            self._join(kv_ready[l])
            self.launches += 1
            ops.attn_fwd(t("q2"), t("k2"), t("v2"), d, d, d, t("o2"), d, t("lse2"), B, Hh, Q, S, scale,
                         **self._attn_drop(f"d{l}_attn2"))
            Wo = D["ca"]["out"]
            self._lin(t("o2"), Wo.Wf, Mq, d, d, d, out=t("pre2"), bias=Wo.bias, residual=t("t1"), ldr=d, **self._drop(f"d{l}_do2"))
            self._ln_fwd(t("pre2"), D["n2"], t("t2"), t("mean2"), t("rstd2"), Mq)
            self._lin(t("t2"), D["l1"].Wf, Mq, dff, d, d, out=t("h"), bias=D["l1"].bias, relu=True, **self._drop(f"d{l}_dh"))
            self._lin(t("h"), D["l2"].Wf, Mq, d, dff, dff, out=t("pre3"), bias=D["l2"].bias, residual=t("t2"), ldr=d,
                      **self._drop(f"d{l}_do3"))
            if l + 1 < self.ndec:
                self._ln_fwd(t("pre3"), D["n3"], t("t3"), t("mean3"), t("rstd3"), Mq, y2=a[f"d{l + 1}_tq"], pos=self.query_pos, S=Q)
            else:
                self._ln_fwd(t("pre3"), D["n3"], t("t3"), t("mean3"), t("rstd3"), Mq)
            tgt = t("t3")
        # shared final norm on every layer output (transformer.py:122-126): one launch over the six layers' rows
        LM = self.ndec * Mq
        self._ln_fwd(a["t3_all"], self.dec_norm, a["hs"], a["meanf"], a["rstdf"], LM)
        # ---------------- heads (detr.py:185-188)
        self._lin(a["hs"], self.h_cls.Wf, LM, self.C, d, d, Cf=a["logits"], ldcf=self.C, bias=self.h_cls.bias)
        self._lin(a["hs"], self.h_b0.Wf, LM, d, d, d, out=a["hb1"], bias=self.h_b0.bias, relu=True)
        self._lin(a["hb1"], self.h_b1.Wf, LM, d, d, d, out=a["hb2"], bias=self.h_b1.bias, relu=True)
        self._lin(a["hb2"], self.h_b2.Wf, LM, 4, d, d, Cf=a["boxes"], ldcf=4, bias=self.h_b2.bias, sigmoid=True)
        self._mark("fwd_decoder_heads")

    def _attn_drop(self, name):
        return dict(self._drop(name), split=self.plane)

    # ------------------------------------------------------------------------------------------ loss
    def match(self, want_cost=False):
        """Hungarian matching of all L*B problems on device (hungarian_matching.py:163-203)."""
        a, L, B, Q = self.a, self.ndec, self.B, self.Q
        cost = None
        if want_cost:
            cost = torch.zeros(L * B, Q, 100, dtype=F32, device=self.device)
        self.launches += 1
        ops.matcher(a["logits"], self.C, a["boxes"], a["t_bbox"], a["t_class"], L * B, B, Q, self.C,
                    a["p_indices"], a["t_indices"], a["p_selector"], a["match"], cost, a["status"])
        return cost

    def set_targets(self, t_bbox, t_class):
        a = self.a
        a["t_bbox"].copy_(t_bbox.reshape(a["t_bbox"].shape), non_blocking=True)
        a["t_class"].copy_(t_class.reshape(a["t_class"].shape), non_blocking=True)

    def loss(self, background_class, loss_scale=1.0, with_grad=True):
        """get_losses (loss.py:22-34): matcher + set criterion for all decoder layers; fills d_logits/d_boxpre."""
        a, L, B, Q = self.a, self.ndec, self.B, self.Q
        self.match()
        self.launches += 3              # clear of the accumulators, loss, finalize
        ops.set_loss(a["logits"], self.C, a["boxes"], a["t_bbox"], a["t_class"], a["match"], L, B, Q, self.C,
                     background_class, self.normalisers, loss_scale, a["loss_sums"], a["losses"], a["total"],
                     a["d_logits"] if with_grad else None, self.ld_dl, a["d_boxpre"] if with_grad else None, 32,
                     status=a["status"], split=self.plane)
        self._mark("matcher_loss")

    def check_matcher_status(self):
        """HOST SYNC.  The reference raises through scipy ("matrix contains invalid numeric entries", hungarian_matching.py:29) when
        a cost matrix holds NaN / -inf; on device that condition sets status != 0 and turns the loss scalars into NaN
        (detrb_set_loss).  Callers that already sync with the host (fit / eval progress prints) turn it back into the exception."""
        if int(self.a["status"].abs().sum()) != 0:
            raise ValueError("matrix contains invalid numeric entries")

    def loss_dict(self, snapshot=False):
        """36 scalars with the reference's keys (loss.py:172-179; aux layer i -> suffix _i, main = last layer).
        The scalars are views of the engine's resident loss buffers, overwritten by the next step; snapshot=True returns views of
        a device copy instead (two small device-to-device copies), safe to read after later steps were enqueued."""
        a, L = self.a, self.ndec
        losses, total = a["losses"], a["total"]
        if snapshot:
            snap = torch.cat((losses.reshape(-1), total))                    # one small device-to-device copy
            if self._distributed() and self.normalisers is not None:
                # data parallel: every rank holds its share of the global-batch loss (local sums / GLOBAL normalisers) -> the
                # logged loss terms are the SUM over ranks; the three accuracy ratios (columns 1-3) are averaged.  One tiny
                # all-reduce on the snapshot; the gradients never depended on it.
                import torch.distributed as dist
                dist.all_reduce(snap, op=dist.ReduceOp.SUM)
                snap[:-1].view(L, 6)[:, 1:4] /= dist.get_world_size()
            losses, total = snap[:-1].view(L, 6), snap[-1:]
        names = ("label_cost", "true_neg", "true_pos", "pos_accuracy", "giou_loss", "l1_loss")
        out = OrderedDict()
        for l in [L - 1] + list(range(L - 1)):
            suf = "" if l == L - 1 else f"_{l}"
            for k, n in enumerate(names):
                out[n + suf] = losses[l, k]
        return total[0], out

    # ------------------------------------------------------------------------------------------ backward
    def zero_grads(self):
        self.grads.zero_()

    def loss_and_zero_grads(self, background_class, loss_scale=1.0):
        """The set loss with its gradient, then the clear of the gradient arena (166 MB).  (The clear BESIDE the matcher on a side
        stream -- the arena is not touched between the previous optimizer step and the first weight gradient -- was measured: no
        difference, 11.42 vs 11.42 ms/step, profiles/r02_zero_grads_fork_ab.log; the extra cross-stream edges cost what the 24 us
        kernel costs.)"""
        self.loss(background_class, loss_scale=loss_scale, with_grad=True)
        self.zero_grads()

    def grad_buckets(self):
        """[lo, hi) ranges of the flat gradient arena in the order the backward pass completes them: transformer + heads
        (+ fine-tuning layers), then layer3-4 + input_proj, then stem + layer1-2.  Data parallel: bucket k's all-reduce runs
        while the backward pass of bucket k+1 is still computing."""
        t0 = self.group_range["transformers"][0]
        l3 = min(o for (name, o, _, _, _) in self.vars if name.startswith("backbone/layer3/"))
        return [(t0, self.total), (l3, t0), (0, l3)]

    def backward(self, train_backbone=True, boundary=None, defer_tail=False):
        """Gradients of total_loss wrt every trainable variable, accumulated (+=) into self.grads.
        train_backbone=False stops after the transformer (results identical for the trained groups; the reference
        computes the dead work anyway, optimizers.py:112-115 / SURVEY appendix A.7).
        boundary(k) is called when gradient bucket k (grad_buckets()) is complete, side-stream weight gradients included.
        defer_tail: return without waiting for the stem's weight gradient (side stream); the optimizer_step() that follows
        applies the other variables beside it and joins.  Single-rank train step only."""
        self._tail = None
        def reached(k):
            if boundary is not None:
                self._join_wgrad()
                boundary(k)
        a, B = self.a, self.B
        d, dff, S, Q, M, Mq, Hh = self.d, self.dff, self.S, self.Q, self.M, self.Mq, self.H
        scale = float(d // Hh) ** -0.5
        LM = self.ndec * Mq
        self._in_backward = True
        inv_keep = 1.0 / (1.0 - self.dropout) if (self.training and self.dropout > 0) else 1.0
        # ---------------- heads
        self._lin_wgrad(self.h_b2, a["hb2"], a["d_boxpre"], LM, ldy=32)
        self._lin(a["d_boxpre"], self.h_b2.Wd, LM, d, 32, 32, out=a["g_hb2"], mask=a["hb2"], ldm=d)
        self._lin_wgrad(self.h_b1, a["hb1"], a["g_hb2"], LM)
        self._lin(a["g_hb2"], self.h_b1.Wd, LM, d, d, d, out=a["g_hb1"], mask=a["hb1"], ldm=d)
        self._lin_wgrad(self.h_b0, a["hs"], a["g_hb1"], LM)
        self._lin(a["g_hb1"], self.h_b0.Wd, LM, d, d, d, out=a["g_hs"])
        self._lin_wgrad(self.h_cls, a["hs"], a["d_logits"], LM, ldy=self.ld_dl)
        self._lin(a["d_logits"], self.h_cls.Wd, LM, d, self.ld_dl, self.ld_dl, out=a["g_hs"], residual=a["g_hs"], ldr=d)
        # ---------------- decoder, last layer first
        # final-norm branch of every layer at once: g_t3f[l] = LN_f'(g_hs[l])
        self._ln_bwd(a["g_hs"], None, a["t3_all"], self.dec_norm, a["meanf"], a["rstdf"], a["g_t3f"], None, LM)
        mem, memp = self.mem, self.memp
        g_next = None                      # gradient wrt t3 coming from layer l+1 (None for the last layer)
        first_mem = True
        for l in reversed(range(self.ndec)):
            D = self.dec[l]
            t = lambda n: a[f"d{l}_{n}"]
            tgt = a["tgt0"] if l == 0 else a[f"d{l - 1}_t3"]
            # LN3 (d t3 = the final-norm branch + the gradient from layer l+1)
            self._ln_bwd(a["g_t3f"][l * Mq:(l + 1) * Mq], g_next, t("pre3"), D["n3"], t("mean3"), t("rstd3"), a["gq_b"], a["gq_c"], Mq, f"d{l}_do3")
            # FFN: pre3 = t2 + drop(h W2 + b2), h = drop(relu(t2 W1 + b1))
            self._lin_wgrad(D["l2"], t("h"), a["gq_c"], Mq)
            self._lin(a["gq_c"], D["l2"].Wd, Mq, dff, d, d, out=a["gq_h"], mask=t("h"), ldm=dff, mask_scale=inv_keep)
            self._lin_wgrad(D["l1"], t("t2"), a["gq_h"], Mq)
            self._lin(a["gq_h"], D["l1"].Wd, Mq, d, dff, dff, out=a["gq_a"], residual=a["gq_b"], ldr=d)      # d t2
            # LN2
            self._ln_bwd(a["gq_a"], None, t("pre2"), D["n2"], t("mean2"), t("rstd2"), a["gq_b"], a["gq_c"], Mq, f"d{l}_do2")
            # cross attention: pre2 = t1 + drop(o2 Wo + bo)
            Wo, W = D["ca"]["out"], D["ca"]["inp"]
            self._lin_wgrad(Wo, t("o2"), a["gq_c"], Mq)
            nd = B * Hh * Q
            dO2, delta2 = a["gq_d2"][l * Mq:(l + 1) * Mq], a["delta2"][l * nd:(l + 1) * nd]      # per layer: read on the side stream
            self._lin(a["gq_c"], Wo.Wd, Mq, d, d, d, out=dO2)                                                  # d o2
            self.launches += 3
            xargs = (t("q2"), t("k2"), t("v2"), t("o2"), dO2, d, d, d, d, d, t("lse2"), delta2,
                     a["gq_q2"], a["gm_k2"], a["gm_v2"], d, d, d, B, Hh, Q, S, scale)
            xkw = self._attn_drop(f"d{l}_attn2")
            self._before_write(a["gq_q2"])
            if self.dec_fork and self.overlap_wgrad:
                # only dQ continues the decoder's chain; dK / dV (1050 keys) feed the weight gradients and d(memory), which
                # live on the side stream anyway: the kernel goes there, in order with its consumers
                ops.attn_bwd(*xargs, parts=1, **xkw)
                self._on_wstream(lambda: ops.attn_bwd(*xargs, parts=2, **xkw), (dO2, delta2), low=True)
                ops.attn_bwd(*xargs, parts=4, **xkw)
            else:
                self._before_write(a["gm_k2"], a["gm_v2"])
                ops.attn_bwd(*xargs, **xkw)
            self._lin_wgrad(W, t("t1q"), a["gq_q2"], Mq, n_off=0, n_rows=d)
            self._lin_wgrad(W, memp, a["gm_k2"], M, n_off=d, n_rows=d, low=True)          # (in order with the dK/dV kernel that feeds them)
            self._lin_wgrad(W, mem, a["gm_v2"], M, n_off=2 * d, n_rows=d, low=True)
            # d memory accumulates over decoder layers (memory feeds every cross attention)
            # (side stream, in order with each other; only the encoder's backward needs g_mem -- joined there)
            def dmem(W=W, first=first_mem):
                self._lin(a["gm_k2"], W.Wd[:, :, d:], M, d, d, W.ldd, out=a["g_mem"], residual=None if first else a["g_mem"], ldr=d)
                self._lin(a["gm_v2"], W.Wd[:, :, 2 * d:], M, d, d, W.ldd, out=a["g_mem"], residual=a["g_mem"], ldr=d)
            if self.overlap_dmem:
                self._on_wstream(dmem, (a["gm_k2"], a["gm_v2"]), low=True)
            else:
                dmem()
            first_mem = False
            self._lin(a["gq_q2"], W.Wd, Mq, d, d, W.ldd, out=a["gq_a"], residual=a["gq_b"], ldr=d)            # d t1
            # LN1
            self._ln_bwd(a["gq_a"], None, t("pre1"), D["n1"], t("mean1"), t("rstd1"), a["gq_b"], a["gq_c"], Mq, f"d{l}_do1")
            # self attention: pre1 = tgt + drop(o Wo + bo)
            Wo, W = D["sa"]["out"], D["sa"]["inp"]
            self._lin_wgrad(Wo, t("o"), a["gq_c"], Mq)
            self._lin(a["gq_c"], Wo.Wd, Mq, d, d, d, out=a["gq_d"])                                             # d o
            self.launches += 3
            # dq | dk | dv are the three column blocks of ONE [rows, 3d] buffer: the gradient wrt the layer input is a single
            # GEMM over K = 3d against the whole in-projection
            gqkv = a["gq_qkv"]
            self._before_write(gqkv, a["delta"])
            sargs = (t("qk"), t("qk")[:, d:], t("v"), t("o"), a["gq_d"], 2 * d, 2 * d, d, d, d, t("lse1"), a["delta"],
                     gqkv, gqkv[:, d:], gqkv[:, 2 * d:], 3 * d, 3 * d, 3 * d, B, Hh, Q, Q, scale)
            skw = self._attn_drop(f"d{l}_attn1")
            if self.dec_fork:                             # dK/dV beside dQ (both short: 100 queries x 100 keys per head)
                ops.attn_bwd(*sargs, parts=1, **skw)
                kvj = self._fork(lambda: ops.attn_bwd(*sargs, parts=2, **skw), lane=1)
                ops.attn_bwd(*sargs, parts=4, **skw)
                self._join(kvj)
            else:
                ops.attn_bwd(*sargs, **skw)
            self._lin_wgrad(W, t("tq"), gqkv, Mq, ldy=3 * d, n_off=0, n_rows=2 * d)
            self._lin_wgrad(W, tgt, gqkv[:, 2 * d:], Mq, ldy=3 * d, n_off=2 * d, n_rows=d)
            if l > 0:
                # d tgt = d_pre1 + dqk.Wqk (tq = tgt + query_pos) + dv.Wv   -> gradient wrt the previous layer's t3
                g_next = a["gq_n0"] if (l & 1) else a["gq_n1"]       # ping-pong: read by layer l-1's LN3 backward
                self._lin(gqkv, W.Wd, Mq, d, 3 * d, W.ldd, out=g_next, residual=a["gq_b"], ldr=d)
        self._mark("bwd_heads_decoder")
        # ---------------- encoder, last layer first.  g_mem = d y2 (last encoder layer output incl. its +pos use)
        self._join_wgrad()                                     # g_mem (and the decoder's weight gradients) complete
        g_y = a["g_mem"]
        for l in reversed(range(self.nenc)):
            E = self.enc[l]
            e = lambda n: a[f"e{l}_{n}"]
            xin = a["src"] if l == 0 else a[f"e{l - 1}_y2"]
            xinp = a["srcp"] if l == 0 else a[f"e{l - 1}_y2p"]
            self._ln_bwd(g_y, None, e("pre2"), E["n2"], e("mean2"), e("rstd2"), a["gm_a"], a["gm_b"], M, f"e{l}_do2")
            self._lin_wgrad(E["l2"], e("h"), a["gm_b"], M)
            self._lin(a["gm_b"], E["l2"].Wd, M, dff, d, d, out=a["gm_h"], mask=e("h"), ldm=dff, mask_scale=inv_keep)
            self._lin_wgrad(E["l1"], e("y1"), a["gm_h"], M)
            self._lin(a["gm_h"], E["l1"].Wd, M, d, dff, dff, out=a["gm_c"], residual=a["gm_a"], ldr=d)        # d y1
            self._ln_bwd(a["gm_c"], None, e("pre1"), E["n1"], e("mean1"), e("rstd1"), a["gm_a"], a["gm_b"], M, f"e{l}_do1")
            Wo, W = E["sa"]["out"], E["sa"]["inp"]
            self._lin_wgrad(Wo, e("o"), a["gm_b"], M)
            self._lin(a["gm_b"], Wo.Wd, M, d, d, d, out=a["gm_c"])                                             # d o
            self.launches += 3
            gqkv = a["gm_qkv"]                                 # dq | dk | dv column blocks of one buffer (see the decoder)
            self._before_write(gqkv, a["delta"])
            self._probed(f"e{l}_attn#bwd", lambda: ops.attn_bwd(
                e("qk"), e("qk")[:, d:], e("v"), e("o"), a["gm_c"], 2 * d, 2 * d, d, d, d, e("lse"), a["delta"],
                gqkv, gqkv[:, d:], gqkv[:, 2 * d:], 3 * d, 3 * d, 3 * d, B, Hh, S, S, scale, **self._attn_drop(f"e{l}_attn")))
            self._lin_wgrad(W, xinp, gqkv, M, ldy=3 * d, n_off=0, n_rows=2 * d)
            self._lin_wgrad(W, xin, gqkv[:, 2 * d:], M, ldy=3 * d, n_off=2 * d, n_rows=d)
            # d x = d_pre1 + [dq | dk | dv] . W_in   (xp = x + pos shares x's gradient): one GEMM over K = 3d
            g_y = a["gm_n0"] if (l & 1) else a["gm_n1"]
            self._lin(gqkv, W.Wd, M, d, 3 * d, W.ldd, out=g_y, residual=a["gm_a"], ldr=d)
        self._mark("bwd_encoder")
        reached(0)
        # ---------------- input_proj
        ip = self.slots["input_proj"]
        self._lin_wgrad(ip, self.feat, g_y, M)
        if not train_backbone:
            self._in_backward = False
            self._join_wgrad()
            reached(1)
            reached(2)
            return
        nb = len(self.blocks)
        last = self.blocks[-1]
        g_out = a["g_x"]
        self._lin(g_y, ip.Wd, M, self.c_feat, d, ip.ldd, out=g_out, mask_bits=a[f"b{nb - 1}_out_bits"], ldmb=self.c_feat // 8)
        # ---------------- backbone, last block first
        g_in = a["g_y"]
        for i in reversed(range(nb)):
            blk = self.blocks[i]
            ihw, ohw = blk["in_hw"], blk["out_hw"]
            x = a["pool"] if i == 0 else a[f"b{i - 1}_out"]
            a1, a2 = a[f"b{i}_a1"], a[f"b{i}_a2"]
            xbits = None if blk["first"] else a[f"b{i - 1}_out_bits"]
            self._conv_wgrad(blk["c3"], a2, g_out, ohw, ohw)
            self._conv_dgrad(blk["c3"], g_out, ohw, ohw, a["g_2"], mask_bits=a[f"b{i}_a2_bits"])
            self._conv_wgrad(blk["c2"], a1, a["g_2"], ihw, ohw)
            self._conv_dgrad(blk["c2"], a["g_2"], ihw, ohw, a["g_1"], mask_bits=a[f"b{i}_a1_bits"])
            self._conv_wgrad(blk["c1"], x, a["g_1"], ihw, ihw)
            if blk["ds"] and blk["stride"] == 1:
                # stride-1 shortcut (layer1 block 0): its data gradient goes first, as a plain GEMM into g_in; conv1's data gradient
                # then adds it IN PLACE as its residual (read tile -> add -> mask -> write the same tile) -- two launches of the fast
                # streaming kernel instead of a read-modify-write scatter epilogue (120 -> ~75 us at batch 8)
                cd = blk["cd"]
                self._conv_wgrad(cd, x, g_out, ihw, ohw)
                Mo = B * ohw[0] * ohw[1]
                self.launches += 1
                self._before_write(g_in)
                ops.igemm(g_out, cd.Wd, Mo, cd.Cin, cd.ldd, cd.N, cd.ldd, ops.plain_geom(Mo, cd.ldd), C=g_in, ldc=cd.Cin,
                          split=self.plane, wsplit=self.wplane)
                self._conv_dgrad(blk["c1"], a["g_1"], ihw, ihw, g_in, mask_bits=xbits, residual=g_in)
            else:
                self._conv_dgrad(blk["c1"], a["g_1"], ihw, ihw, g_in, mask_bits=xbits, residual=None if blk["ds"] else g_out)
            if blk["ds"] and blk["stride"] != 1:
                cd = blk["cd"]
                self._conv_wgrad(cd, x, g_out, ihw, ohw)
                st = blk["stride"]
                Mo = B * ohw[0] * ohw[1]
                g = dict(batch=B, IH=ohw[0], IW=ohw[1], Cin=cd.ldd, OH=ohw[0], OW=ohw[1], KH=1, KW=1, stride=1, pad=0, mode=0)
                self.launches += 1
                self._before_write(g_in)
                ops.igemm(g_out, cd.Wd, Mo, cd.Cin, cd.ldd, cd.N, cd.ldd, g, mask_bits=xbits, ldmb=cd.Cin // 8, mask_scale=1.0,
                          C=g_in, ldc=cd.Cin, out_stride=st, SH=ihw[0], SW=ihw[1], accumulate=True, split=self.plane, wsplit=self.wplane,
                          scratch=a.get("g_s"))
            g_out, g_in = g_in, g_out
            if blk["prefix"] == "backbone/layer3/0":
                reached(1)
        # g_out now holds d pool
        self.launches += 1
        self._before_write(g_in)
        HP, WP = self.hw_pad
        stem = self.slots["backbone/conv1"]
        self.launches += 1
        x, Ms = a["s2d"], self.M_stem
        before_stem = self._w_last if self._w_last is not None else True
        pool_arg = a["pool_arg"]

        def pool_and_stem():
            ops.maxpool_bwd(g_out, pool_arg, g_in, B, self.hw_stem[0], self.hw_stem[1], 64, self.hw_pool[0], self.hw_pool[1],
                            XH=HP, XW=WP, split=self.plane)
            ops.wgrad(x, 16, g_in, stem.N, Ms, stem.N, stem.K, ops.plain_geom(Ms, stem.K), stem.grad, stem.K,
                      rowscale=stem.fold, dbias=stem.bias_grad, a_kb_rows=WP, k_mask=True, split=self.plane)
        # the max-pool backward feeds nothing but the stem's weight gradient.  DETRB_POOL_BWD_SIDE=1 sends both to the side stream, so
        # that the optimizer step of everything else (optimizer_step, defer_tail) starts beside the pooling kernel instead of behind
        # it: measured, no difference (11.42 vs 11.43 ms/step, profiles/r02_pool_bwd_side_ab.log) -- pooling, stem weight gradient
        # and Adam are all HBM-bound, their sum is what the tail costs either way -- so the pooling kernel stays on the main stream
        if os.environ.get("DETRB_POOL_BWD_SIDE") != "1":
            ops.maxpool_bwd(g_out, pool_arg, g_in, B, self.hw_stem[0], self.hw_stem[1], 64, self.hw_pool[0], self.hw_pool[1],
                            XH=HP, XW=WP, split=self.plane)
            self._on_wstream(lambda: ops.wgrad(x, 16, g_in, stem.N, Ms, stem.N, stem.K, ops.plain_geom(Ms, stem.K), stem.grad, stem.K,
                                               rowscale=stem.fold, dbias=stem.bias_grad, a_kb_rows=WP, k_mask=True, split=self.plane), (x, g_in))
        else:
            self._on_wstream(pool_and_stem, (x, g_in, g_out, pool_arg))
        self._in_backward = False
        if defer_tail and boundary is None and self._w_last is not None and hasattr(self.lib, "detrb_adam_clipnorm_chunked"):
            self._tail = before_stem                               # joined by optimizer_step()
            self._mark("bwd_backbone")
            return
        self._join_wgrad()
        reached(2)
        self._mark("bwd_backbone")

    # ------------------------------------------------------------------------------------------ optimizer
    def set_lrs(self, backbone_lr, transformers_lr, nlayers_lr=0.0):
        key = (float(backbone_lr), float(transformers_lr), float(nlayers_lr))
        if getattr(self, "_lrs_host", None) != key:             # (a host->device copy from pageable memory blocks the host)
            self.lrs[:3] = torch.tensor(key, dtype=F32)
            self._lrs_host = key

    def set_enabled(self, backbone, transformers, nlayers=False):
        key = (int(bool(backbone)), int(bool(transformers)), int(bool(nlayers)))
        if getattr(self, "_enabled_host", None) != key:        # (a host->device copy from pageable memory blocks the host)
            self.group_enabled[:3] = torch.tensor(key, dtype=torch.uint8)
            self._enabled_host = key

    def fused_step(self, background_class, clipnorm):
        """forward + losses + backward + gradient all-reduce + Adam for every enabled group + weight refresh on the staged batch
        as ONE replayed launch sequence (capture_train_step): what training.fit runs when there is no gradient accumulation.  The
        capture does not execute anything, so the caller must have run one step through grads_step / apply_group before (kernel
        attribute set-up, NCCL communicator); group flags and learning rates are read from device memory at replay time."""
        key = (self.plan_key, int(background_class), float(clipnorm), self.normalisers is not None, self.u8_input,
               getattr(self, "input_method", None), self._distributed(), getattr(self, "s2d_staged", False))
        if getattr(self, "_fs_key", None) != key:
            self._fs_replay = self.capture_train_step(background_class, clipnorm, warmup=0)
            self._fs_key = key
        self._fs_replay()

    def apply_group(self, name, grads_arena, clipnorm):
        """Adam apply for ONE group (the reference applies its three optimizers one after the other,
        training.py:53-54 -> optimizers.py:160-163): only the group's chunks of the table are visited; the bf16 kernel-layout
        weight copies are refreshed once, before the next forward pass (_ensure_weights), not once per group."""
        gi = GROUPS.index(name)
        if getattr(self, "_en_one", None) is None:
            self._en_one = [torch.zeros(8, dtype=torch.uint8).index_fill_(0, torch.tensor([i]), 1).to(self.device) for i in range(len(GROUPS))]
        lo, hi = self.group_chunks[name]
        self.launches += 3
        # its own one-hot flag vector: self.group_enabled (read at replay time by the captured whole-step graph) stays untouched
        self._adam(grads_arena, clipnorm, lo=lo, hi=hi, enabled=self._en_one[gi])
        self._weights_dirty = True

    def _ensure_weights(self):
        if getattr(self, "_weights_dirty", False):
            self._weights_dirty = False
            self.refresh_weights()

    def _adam(self, grads_arena, clipnorm, lo=0, hi=None, prologue=True, enabled=None):
        """Adam + per-variable clipnorm over chunks [lo, hi) of the chunk table (whole variables); prologue: first call of the step"""
        enabled = self.group_enabled if enabled is None else enabled
        if hasattr(self.lib, "detrb_adam_clipnorm_chunked"):
            hi = self.nchunks if hi is None else hi
            ops.adam_clipnorm_chunked(self.params, grads_arena, self.adam_m, self.adam_v, self.chunks, hi - lo, self.lr_group,
                                      self.lrs, enabled, self.T, clipnorm, self.steps, self.norms, first_chunk=lo,
                                      prologue=prologue)
        else:                                                    # table variant: every variable, the group flags select
            ops.adam_clipnorm(self.params, grads_arena, self.adam_m, self.adam_v, self.table, self.lr_group, self.lrs,
                              enabled, self.T, self.total, clipnorm, self.steps, self.norms)

    def allreduce_grads(self):
        """data parallel: ONE sum all-reduce over the flat gradient arena (NCCL over NVLink); no-op on one rank."""
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.grads, op=dist.ReduceOp.SUM)

    def allreduce_bucket(self, k):
        """data parallel: asynchronous sum all-reduce of gradient bucket k (grad_buckets()); ordered after the work enqueued on
        the current stream, runs on the process group's own stream beside the rest of the backward pass.  Returns the Work
        handle -- wait() on it before the optimizer reads the gradients."""
        import torch.distributed as dist
        lo, hi = self.grad_buckets()[k]
        return dist.all_reduce(self.grads[lo:hi], op=dist.ReduceOp.SUM, async_op=True)

    def set_global_normalisers(self, t_bbox):
        """Under data parallelism the reference's batch-level normalisers (loss.py:66-67,82,94) must be those of the
        GLOBAL batch: N = sum n_i over all ranks, sum_w = 0.1*(B_glob*Q - N) + N.  They depend on the labels only, so
        one tiny all-reduce of the local target count before the step is enough."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
            self.normalisers = None
            return
        n_local = t_bbox.reshape(-1, 100, 4)[:, 0, 0].to(self.device, torch.float32).clamp(0, 99).sum().reshape(1)
        dist.all_reduce(n_local, op=dist.ReduceOp.SUM)
        bq = float(dist.get_world_size() * t_bbox.shape[0] * self.Q)
        if getattr(self, "_norm_buf", None) is None:
            self._norm_buf = torch.zeros(2, dtype=F32, device=self.device)
        self._norm_buf[0:1] = n_local
        self._norm_buf[1:2] = 0.1 * (bq - n_local) + n_local
        self.normalisers = self._norm_buf

    def optimizer_step(self, clipnorm):
        """aggregate_grad_and_apply's apply branch (optimizers.py:160-163) for all enabled groups.  After backward(defer_tail=True)
        the stem's weight gradient -- the last kernel of the backward pass, 155 us alone on the GPU -- is still running on the side
        stream: every other variable is applied beside it, the stem kernel (the first chunks of the table) afterwards."""
        if getattr(self, "_tail", None) is not None:
            before_stem, self._tail = self._tail, None
            if before_stem is not True:
                torch.cuda.current_stream().wait_event(before_stem)       # all weight gradients issued before the stem's
            self.launches += 5
            self._adam(self.grads, clipnorm, lo=self.stem_chunks, prologue=True)
            self._join_wgrad()
            self._adam(self.grads, clipnorm, lo=0, hi=self.stem_chunks, prologue=False)
        else:
            self.launches += 3
            self._adam(self.grads, clipnorm)
        self.refresh_weights()

    # ------------------------------------------------------------------------------------------ fused fast path
    def train_step(self, background_class, clipnorm, loss_scale=1.0, train_backbone=True):
        """One optimizer step on the batch already resident in a['images'] / a['t_bbox'] / a['t_class']:
        training.py:9-25 + :53-54 with no accumulation.  Launch-only (no host sync): CUDA-graph capturable."""
        self.training = True
        self._ensure_weights()
        self.seed_dev.add_(1)                   # fresh dropout masks every step, also under graph replay
        self._forward_impl()
        self.loss_and_zero_grads(background_class, loss_scale)
        if self._distributed():
            works = []
            self.backward(train_backbone=train_backbone, boundary=lambda k: works.append(self.allreduce_bucket(k)))
            for w in works:
                w.wait()
        else:
            self.backward(train_backbone=train_backbone, defer_tail=True)
        self._mark("allreduce")
        self.optimizer_step(clipnorm)
        self._mark("adam_refresh")

    def _distributed(self):
        import torch.distributed as dist
        return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def capture_train_step(self, background_class, clipnorm, loss_scale=1.0, train_backbone=True, warmup=2):
        """Capture train_step into CUDA graphs (static shapes: fixed-size images).  Returns a callable replaying it.
        Single rank: one graph for the whole step.  Data parallel: the step is cut at the gradient-bucket boundaries
        (grad_buckets()) into four graphs; the all-reduces stay eager NCCL calls between them (NCCL's watchdog threads and
        graph capture do not mix safely), asynchronous, so bucket k is reduced over NVLink while graph k+1 computes."""
        def part1():
            self.training = True
            self.seed_dev.add_(1)
            self._forward_impl()
            self.loss_and_zero_grads(background_class, loss_scale)
            self.backward(train_backbone=train_backbone, defer_tail=not self._distributed())

        def part2():
            self.optimizer_step(clipnorm)
        self._ensure_weights()
        dist_on = self._distributed()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                part1()
                self.allreduce_grads()
                part2()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = self.launches
        if not dist_on:
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph, stream=self._chain_stream()):
                part1()
                part2()
            self.launches_per_step = self.launches - n0
            self._graph = graph
            return graph.replay
        # (One graph with the bucket all-reduces captured inside it was tried once more at the end of round 2 -- torch 2.11 / NCCL 2.28,
        #  two B200s: the capture completes, the first replay never returns -- a hang, killed by the test's timeout.  The cut stays.)
        # data parallel: the backward pass is cut at the gradient-bucket boundaries into consecutive graphs; bucket k's eager
        # all-reduce is enqueued between graph k and graph k+1 and runs (on NCCL's stream) while graph k+1 computes
        import gc
        graphs = []
        pool = torch.cuda.graph_pool_handle()
        cap = self._chain_stream()

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool, capture_error_mode="thread_local")
            graphs.append(g)

        def cut(k):
            graphs[-1].capture_end()
            begin()
        torch.cuda.synchronize()
        gc.collect()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            begin()
            self.training = True
            self.seed_dev.add_(1)
            self._forward_impl()
            self.loss_and_zero_grads(background_class, loss_scale)
            self.backward(train_backbone=train_backbone, boundary=cut)      # graphs 0..2 end at the bucket boundaries
            part2()                                                           # graph 3: optimizer
            graphs[-1].capture_end()
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        self.launches_per_step = self.launches - n0
        self._graph = tuple(graphs)
        assert len(graphs) == 4, len(graphs)

        overlap = os.environ.get("DETRB_DP_OVERLAP", "1") != "0"      # 0: one flat all-reduce after the backward pass (A/B runs)

        def replay():
            works = []
            for k in range(3):
                graphs[k].replay()
                if overlap:
                    works.append(self.allreduce_bucket(k))
            if not overlap:
                self.allreduce_grads()
            for w in works:
                w.wait()
            graphs[3].replay()
        return replay

    # ------------------------------------------------------------------------------------------ graph-replayed gradient step
    def stage_inputs(self, images, t_bbox, t_class, direct=False):
        """host/device inputs -> the engine's resident buffers (async copies on the current stream; direct: see _stage_images)"""
        B, H, W, _ = images.shape
        self._plan(B, H, W)
        self._stage_images(images, direct=direct)
        self.set_targets(t_bbox, t_class)

    def _capture_bucketed(self, body):
        """Capture body(boundary) -- a backward pass that calls boundary(k) when gradient bucket k is complete -- as consecutive
        CUDA graphs cut at the bucket boundaries (three graphs for grad_buckets()).  NCCL calls stay out of the graphs (NCCL's
        watchdog threads and graph capture do not mix safely): the caller replays graph k, enqueues bucket k's asynchronous
        all-reduce, replays graph k+1, ... so that bucket k crosses NVLink while graph k+1 computes."""
        import gc
        graphs = []
        pool = torch.cuda.graph_pool_handle()
        cap = self._chain_stream()

        def begin():
            g = torch.cuda.CUDAGraph()
            g.capture_begin(pool=pool, capture_error_mode="thread_local")
            graphs.append(g)
        nb = len(self.grad_buckets())

        def cut(k):
            graphs[-1].capture_end()
            if k + 1 < nb:
                begin()
        torch.cuda.synchronize()
        gc.collect()
        cap.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(cap):
            begin()
            body(cut)
        torch.cuda.current_stream().wait_stream(cap)
        torch.cuda.synchronize()
        assert len(graphs) == nb, len(graphs)
        return graphs

    def grads_step(self, background_class, loss_scale=1.0, use_graph=True):
        """forward(training=True) + set loss + zero_grads + backward + gradient all-reduce on the staged batch
        (training.py:9-25).  With use_graph the launch sequence is captured once per (shape, arguments) and replayed.
        Data parallel: the step is cut at the gradient-bucket boundaries (grad_buckets()); bucket k's sum all-reduce is enqueued
        asynchronously as soon as its part of the backward pass has been replayed and runs over NVLink while the next part
        computes; the current stream waits for all of them before returning (stream-level wait, no host sync)."""
        def body(boundary=None):
            self.training = True
            self.seed_dev.add_(1)
            self._forward_impl()
            self.loss_and_zero_grads(background_class, loss_scale)
            self.backward(train_backbone=True, boundary=boundary)
        self._ensure_weights()                             # a per-group apply (apply_group) since the last forward pass
        dist_on = self._distributed()
        if not use_graph or self.device.type != "cuda":
            if not dist_on:
                return body()
            works = []
            body(lambda k: works.append(self.allreduce_bucket(k)))
            for w in works:
                w.wait()
            return
        key = (self.plan_key, int(background_class), float(loss_scale), self.normalisers is not None, self.u8_input,
               getattr(self, "input_method", None), dist_on, getattr(self, "s2d_staged", False))
        if getattr(self, "_gs_key", None) != key:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body()                                     # warm-up (kernel attribute setup)
                self.allreduce_grads()                     # (and NCCL's communicator, outside any capture)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            if dist_on:
                self._gs_graph = self._capture_bucketed(body)
            else:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._chain_stream(), capture_error_mode="thread_local"):
                    body()
                self._gs_graph = g
            self._gs_key = key
        if not dist_on:
            return self._gs_graph.replay()
        works = []
        for k, g in enumerate(self._gs_graph):
            g.replay()
            works.append(self.allreduce_bucket(k))         # eager, asynchronous NCCL call between two graphs
        for w in works:
            w.wait()
