"""Checkpoint I/O.

The reference's networks/weights.py:5-36 downloads a TF object-graph checkpoint ("detr.ckpt") from a bucket and calls
Keras load_weights -- neither the network nor TensorFlow exist here.  What a user of the hot path needs instead:

* save_checkpoint / load_checkpoint: parameters (reference layouts: HWIO convs, [out,in] Linear, [in,out] Dense), Adam
  moments and iteration counts in one .npz -- resume is absent from the reference (fit() never saves);
* from_torch_detr_state_dict: the ORIGINAL DETR release (facebookresearch/detr `detr-r50-e632da11.pth`, the weights the
  reference's checkpoint was converted from: custom_layers.py:32-35 / transformer.py:250-268 keep torch's layouts) ->
  this package's names.  Conv kernels OIHW -> HWIO; everything else is copied as is;
* load_weights(model, weights): the reference's entry point; accepts a path to either file kind.
"""
import os
from collections import OrderedDict

import numpy as np
import torch

from .spec import RESNET_STAGES, model_params


def save_checkpoint(model, path, config=None):
    eng = model.engine
    state = {k: (v.numpy() if isinstance(v, torch.Tensor) else v) for k, v in eng.export_state().items()}
    if config is not None:
        state["global_step"] = np.int64(getattr(config, "global_step", 0))
    tmp = path + ".tmp.npz"
    np.savez(tmp, **state)
    os.replace(tmp, path)              # atomic: a crash while writing never leaves a truncated checkpoint behind
    return path


def load_checkpoint(model, path, config=None):
    with np.load(path) as z:
        state = {k: z[k] for k in z.files}
    model.engine.load_state(state)
    if config is not None and "global_step" in state:
        config.global_step = int(state["global_step"])
    return model


def _torch_name_map(backbone="resnet50", num_encoder_layers=6, num_decoder_layers=6, nb_class=None):
    """{our name: (torch state_dict key, is_conv)} for the original DETR module tree"""
    m = OrderedDict()

    def bn(ours, theirs):
        for suf in ("weight", "bias", "running_mean", "running_var"):
            m[f"{ours}/{suf}"] = (f"{theirs}.{suf}", False)
    body = "backbone.0.body"
    m["backbone/conv1/kernel"] = (f"{body}.conv1.weight", True)
    bn("backbone/bn1", f"{body}.bn1")
    for li, (nb, _, _, _) in enumerate(RESNET_STAGES[backbone]):
        for b in range(nb):
            ours, theirs = f"backbone/layer{li + 1}/{b}", f"{body}.layer{li + 1}.{b}"
            for c in (1, 2, 3):
                m[f"{ours}/conv{c}/kernel"] = (f"{theirs}.conv{c}.weight", True)
                bn(f"{ours}/bn{c}", f"{theirs}.bn{c}")
            if b == 0:
                m[f"{ours}/downsample_0/kernel"] = (f"{theirs}.downsample.0.weight", True)
                bn(f"{ours}/downsample_1", f"{theirs}.downsample.1")
    m["input_proj/kernel"] = ("input_proj.weight", True)
    m["input_proj/bias"] = ("input_proj.bias", False)

    def mha(ours, theirs):
        m[f"{ours}/in_proj_kernel"] = (f"{theirs}.in_proj_weight", False)
        m[f"{ours}/in_proj_bias"] = (f"{theirs}.in_proj_bias", False)
        m[f"{ours}/out_proj_kernel"] = (f"{theirs}.out_proj.weight", False)
        m[f"{ours}/out_proj_bias"] = (f"{theirs}.out_proj.bias", False)

    def lin(ours, theirs):
        m[f"{ours}/kernel"] = (f"{theirs}.weight", False)
        m[f"{ours}/bias"] = (f"{theirs}.bias", False)

    def ln(ours, theirs):
        m[f"{ours}/gamma"] = (f"{theirs}.weight", False)
        m[f"{ours}/beta"] = (f"{theirs}.bias", False)
    for l in range(num_encoder_layers):
        ours, theirs = f"transformer/encoder/layer_{l}", f"transformer.encoder.layers.{l}"
        mha(ours + "/self_attn", theirs + ".self_attn")
        lin(ours + "/linear1", theirs + ".linear1")
        lin(ours + "/linear2", theirs + ".linear2")
        ln(ours + "/norm1", theirs + ".norm1")
        ln(ours + "/norm2", theirs + ".norm2")
    for l in range(num_decoder_layers):
        ours, theirs = f"transformer/decoder/layer_{l}", f"transformer.decoder.layers.{l}"
        mha(ours + "/self_attn", theirs + ".self_attn")
        mha(ours + "/multihead_attn", theirs + ".multihead_attn")
        lin(ours + "/linear1", theirs + ".linear1")
        lin(ours + "/linear2", theirs + ".linear2")
        for k in (1, 2, 3):
            ln(ours + f"/norm{k}", theirs + f".norm{k}")
    ln("transformer/decoder/norm", "transformer.decoder.norm")
    m["query_embed/kernel"] = ("query_embed.weight", False)
    if nb_class is None:
        lin("class_embed", "class_embed")
        for k in range(3):
            lin(f"bbox_embed_{k}", f"bbox_embed.layers.{k}")
    return m


def from_torch_detr_state_dict(sd, backbone="resnet50", num_encoder_layers=6, num_decoder_layers=6, nb_class=None,
                               fill=None):
    """Original-DETR state_dict -> {our name: tensor in the reference layout}.  With nb_class the new fine-tuning heads
    (absent from the checkpoint, detr.py:94-114) are taken from `fill` (e.g. init_params(nb_class=...))."""
    if "model" in sd and isinstance(sd["model"], dict):
        sd = sd["model"]                                   # the released .pth wraps the weights as {"model": state_dict}
    spec = model_params(backbone=backbone, num_encoder_layers=num_encoder_layers, num_decoder_layers=num_decoder_layers,
                        nb_class=nb_class)
    names = _torch_name_map(backbone, num_encoder_layers, num_decoder_layers, nb_class)
    out = OrderedDict()
    for name, p in spec.items():
        if name in names:
            key, is_conv = names[name]
            if key not in sd:
                raise KeyError(f"{key} (for {name}) missing from the state_dict")
            t = torch.as_tensor(sd[key]).detach().to(torch.float32)
            if is_conv:
                t = t.permute(2, 3, 1, 0).contiguous()                   # OIHW -> HWIO (Keras Conv2D)
        else:
            if fill is None or name not in fill:
                raise KeyError(f"{name} is not in the original checkpoint: pass fill= with the new heads")
            t = torch.as_tensor(fill[name]).to(torch.float32)
        if tuple(t.shape) != p.shape:
            raise ValueError(f"{name}: shape {tuple(t.shape)} != {p.shape}")
        out[name] = t
    return out


def load_weights(model, weights: str):
    """networks/weights.py:14-36.  `weights`: path of a checkpoint written by save_checkpoint (.npz) or of an original
    DETR .pth; the reference's named download ("detr") needs a network and a TF checkpoint reader -- not available."""
    eng = model.engine
    if isinstance(weights, str) and weights.endswith(".npz") and os.path.exists(weights):
        return load_checkpoint(model, weights)
    if isinstance(weights, str) and weights.endswith((".pth", ".pt")) and os.path.exists(weights):
        sd = torch.load(weights, map_location="cpu", weights_only=True)
        fill = None
        if eng.nb_class is not None:
            fill = model.export_params()                   # keep the freshly initialised fine-tuning heads
        model.load_params(from_torch_detr_state_dict(sd, eng.backbone_name, eng.nenc, eng.ndec, eng.nb_class, fill))
        return model
    raise Exception(f"Cant load the weights: {weights} (the reference's bucket download is not reachable offline; "
                    f"pass a .npz written by save_checkpoint or the original DETR .pth)")
