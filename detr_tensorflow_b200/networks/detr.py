"""get_detr_model -- mirror of detr_tf/networks/detr.py:116-204 on top of the sm_100a engine."""
import numpy as np
import torch

from ..engine import Engine
from .init import init_params


def _as_device_f32(x, device):
    """float images go in as fp32 NHWC (the reference's input); uint8 frames stay uint8: the engine normalises them on device"""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if x.dtype == torch.uint8:
        return x.to(device=device, non_blocking=True)
    return x.to(device=device, dtype=torch.float32, non_blocking=True)


class DetrModel:
    """Callable like the Keras functional model the reference returns: model(images[B,H,W,3], training=bool) ->
    {'pred_logits': [B,100,C], 'pred_boxes': [B,100,4], 'aux': [5 x same]} (detr.py:190-204).
    With include_top=False the call returns the stacked decoder states hs [L,B,100,256] (detr.py:177-179)."""

    def __init__(self, engine, include_top, name, config=None):
        self.engine = engine
        self.include_top = include_top
        self.name = name
        self.config = config

    def __call__(self, images, training=False):
        """images: [B,H,W,3] float32, already normalised (the reference's input) -- or, extension, raw uint8 frames, which are
        normalised on device per config.normalized_method (data/processing.py:6-23) while being laid out for the stem."""
        eng = self.engine
        images = _as_device_f32(images, eng.device)
        if images.dtype == torch.uint8:
            method = getattr(self.config, "normalized_method", "torch_resnet")
            if getattr(eng, "input_method", None) != method:
                eng.set_input_normalisation(method)
        out = eng.forward(images, training=training)
        if self.include_top:
            return out
        return eng.a["hs"].view(eng.ndec, eng.B, eng.Q, eng.d)

    # convenience (not in the reference): parameter I/O in the reference's layouts
    def load_params(self, params):
        self.engine.load_params(params)

    def export_params(self):
        return self.engine.export_params()


def get_detr_model(config, include_top=False, nb_class=None, weights=None, tf_backbone=False, num_decoder_layers=6,
                   num_encoder_layers=6, backbone="resnet50", device="cuda", seed=0, params=None, dropout=0.1, precision="bf16"):
    """detr.py:116-204.  Extensions: `backbone` ("resnet50" | "resnet101": the reference ignores its own backbone
    argument, detr.py:21,31), `device`, `seed`/`params` (synthetic or caller-provided weights in reference layouts),
    `dropout` (transformer.py:9 default 0.1; 0 makes training-mode steps deterministic for tests), `precision`: "bf16"
    (throughput: bf16 activations, one tensor-core pass) or "parity" (activations and weight copies as bf16 PAIRS = 16
    significant bits, three tensor-core passes per product, fp32 attention: the reference's fp32 arithmetic to ~1e-4)."""
    if tf_backbone:
        raise NotImplementedError("tf_backbone=True (keras.applications ResNet50, detr.py:146-148) is out of scope")
    # detr.py:178-181: nb_class only matters when include_top is False -> add_heads_nlayers (detr.py:94-114): new Keras
    # Dense heads `cls_layer` (nb_class logits) and `pos_layer` (256-256-4 MLP), registered as config.nlayers (detr.py:103)
    finetune = (include_top is False) and (nb_class is not None)
    eng = Engine(device=device, backbone=backbone, num_classes=92, num_encoder_layers=num_encoder_layers,
                 num_decoder_layers=num_decoder_layers, seed=seed, dropout=dropout, nb_class=nb_class if finetune else None,
                 precision=precision)
    if finetune:
        config.add_nlayers(["cls_layer", "pos_layer"])
    if params is None:
        params = init_params(seed, backbone=backbone, num_encoder_layers=num_encoder_layers,
                             num_decoder_layers=num_decoder_layers, nb_class=nb_class if finetune else None)
    eng.load_params(params)
    has_heads = bool(include_top) or finetune
    model = DetrModel(eng, has_heads, "detr_finetuning" if has_heads else "detr", config)
    if weights is not None:                               # detr.py:143-144 -> networks/weights.py:14 (local files only)
        from .weights import load_weights
        load_weights(model, weights)
    return model
