"""Parameter inventory of the DETR model in the reference's naming / layouts.

Mirrors what `get_detr_model(include_top=True)` owns (networks/detr.py:19-53, resnet_backbone.py:7-136,
transformer.py:7-268, custom_layers.py): Keras Conv2D kernels are HWIO, `Linear` kernels are [out, in], the
attention in-projection is one packed [3d, d] tensor.  `group` follows optimizers.py:10-43.
"""
from collections import OrderedDict

RESNET_STAGES = {
    "resnet50": [(3, 64, 256, 1), (4, 128, 512, 2), (6, 256, 1024, 2), (3, 512, 2048, 2)],
    "resnet101": [(3, 64, 256, 1), (4, 128, 512, 2), (23, 256, 1024, 2), (3, 512, 2048, 2)],
}


class P:
    __slots__ = ("name", "shape", "kind", "group")

    def __init__(self, name, shape, kind, group):
        self.name, self.shape, self.kind, self.group = name, tuple(shape), kind, group


def model_params(num_classes=92, backbone="resnet50", num_encoder_layers=6, num_decoder_layers=6,
                 model_dim=256, ffn_dim=2048, num_queries=100, nb_class=None):
    """OrderedDict name -> P.  kind: conv | bn_w | bn_b | bn_mean | bn_var | linear_w | dense_w | linear_b | ln_g | ln_b |
    embed.  group: 'backbone' | 'transformers' | 'nlayers' | None (non-trainable / in no optimizer group).
    nb_class: the fine-tuning model of add_heads_nlayers (detr.py:94-114) -- `cls_layer` / `pos_layer` Keras Dense heads
    (kernel [in, out]) in the 'nlayers' group (optimizers.py:39-43) instead of class_embed / bbox_embed_*."""
    out = OrderedDict()

    def add(name, shape, kind, group):
        out[name] = P(name, shape, kind, group)

    def bn(prefix, c):
        for suf, kind in (("weight", "bn_w"), ("bias", "bn_b"), ("running_mean", "bn_mean"), ("running_var", "bn_var")):
            add(f"{prefix}/{suf}", (c,), kind, None)       # trainable=False, custom_layers.py:11-18

    add("backbone/conv1/kernel", (7, 7, 3, 64), "conv", "backbone")
    bn("backbone/bn1", 64)
    cin = 64
    for li, (nb, d1, d2, stride) in enumerate(RESNET_STAGES[backbone]):
        for b in range(nb):
            p = f"backbone/layer{li + 1}/{b}"
            add(p + "/conv1/kernel", (1, 1, cin, d1), "conv", "backbone")
            bn(p + "/bn1", d1)
            add(p + "/conv2/kernel", (3, 3, d1, d1), "conv", "backbone")
            bn(p + "/bn2", d1)
            add(p + "/conv3/kernel", (1, 1, d1, d2), "conv", "backbone")
            bn(p + "/bn3", d2)
            if b == 0:
                add(p + "/downsample_0/kernel", (1, 1, cin, d2), "conv", "backbone")
                bn(p + "/downsample_1", d2)
            cin = d2
    d = model_dim
    # input_proj belongs to the "backbone" optimizer group: every layer of the inner model except
    # `transformer` (optimizers.py:25-36)
    add("input_proj/kernel", (1, 1, cin, d), "conv", "backbone")
    add("input_proj/bias", (d,), "linear_b", "backbone")

    def mha(p):
        add(p + "/in_proj_kernel", (3 * d, d), "linear_w", "transformers")
        add(p + "/in_proj_bias", (3 * d,), "linear_b", "transformers")
        add(p + "/out_proj_kernel", (d, d), "linear_w", "transformers")
        add(p + "/out_proj_bias", (d,), "linear_b", "transformers")

    def lin(p, o, i):
        add(p + "/kernel", (o, i), "linear_w", "transformers")
        add(p + "/bias", (o,), "linear_b", "transformers")

    def ln(p):
        add(p + "/gamma", (d,), "ln_g", "transformers")
        add(p + "/beta", (d,), "ln_b", "transformers")

    for l in range(num_encoder_layers):
        p = f"transformer/encoder/layer_{l}"
        mha(p + "/self_attn")
        lin(p + "/linear1", ffn_dim, d)
        lin(p + "/linear2", d, ffn_dim)
        ln(p + "/norm1")
        ln(p + "/norm2")
    for l in range(num_decoder_layers):
        p = f"transformer/decoder/layer_{l}"
        mha(p + "/self_attn")
        mha(p + "/multihead_attn")
        lin(p + "/linear1", ffn_dim, d)
        lin(p + "/linear2", d, ffn_dim)
        ln(p + "/norm1")
        ln(p + "/norm2")
        ln(p + "/norm3")
    ln("transformer/decoder/norm")
    # query_embed(None) is evaluated outside the Keras graph: in no optimizer group
    add("query_embed/kernel", (num_queries, d), "embed", None)
    if nb_class is None:
        lin("class_embed", num_classes, d)
        lin("bbox_embed_0", d, d)
        lin("bbox_embed_1", d, d)
        lin("bbox_embed_2", 4, d)
    else:
        for p, i, o in (("cls_layer", d, nb_class), ("pos_layer/dense", d, d), ("pos_layer/dense_1", d, d),
                        ("pos_layer/dense_2", d, 4)):
            add(p + "/kernel", (i, o), "dense_w", "nlayers")
            add(p + "/bias", (o,), "linear_b", "nlayers")
    return out


def head_names(nb_class=None):
    """(class head, box MLP layer 0, 1, 2) slot prefixes"""
    if nb_class is None:
        return "class_embed", "bbox_embed_0", "bbox_embed_1", "bbox_embed_2"
    return "cls_layer", "pos_layer/dense", "pos_layer/dense_1", "pos_layer/dense_2"
