"""Generate golden vectors for the MODEL path (backbone -> input_proj -> position embedding -> transformer -> heads) by
EXECUTING THE REFERENCE'S OWN CODE.

    python tests/golden/make_golden_model.py            # needs /root/reference (build container only)

detr_tf/networks/{detr,resnet_backbone,transformer,custom_layers,position_embeddings}.py are imported unmodified from
/root/reference with `tensorflow` replaced by tests/golden/tf_torch_model_shim.py (TF itself is not installable here):
`get_detr_model(config, include_top=True, ...)` (detr.py:116-204) builds the model and runs it on a seeded image with the
oracle's seeded weights injected by variable name.  Outputs -> tests/golden/model_golden.npz (committed; /root/reference does
not exist on the GPU box).  The weights are NOT stored: tests regenerate them from the same seed (oracle.init_params).
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import tf_torch_model_shim as shim  # noqa: E402

shim.install()
sys.path.insert(0, "/root/reference")
from detr_tf.networks import detr as ref_detr  # noqa: E402
from oracle import detr_oracle as O  # noqa: E402

CASES = {                # name: (seed, batch, H, W, encoder layers, decoder layers, nb_class)
    "a": (11, 2, 96, 128, 2, 3, None),
    "b": (12, 1, 75, 110, 1, 2, None),     # odd sizes: every stride-2 stage rounds
    "ft": (13, 1, 64, 96, 1, 6, 3),        # include_top=False, nb_class=3: add_heads_nlayers (detr.py:94-114; 5 aux outputs hard-coded)
}


class Cfg:
    normalized_method = "torch_resnet"
    nlayers = []

    def add_nlayers(self, layers):           # training_config.py:79-82
        self.nlayers = [l.name for l in layers]


TRAIN = (21, 2, 96, 128, 1, 2, 4)          # seed, batch, H, W, encoder layers, decoder layers, targets per image


def train_case():
    """Gradient of the reference's own loss through the reference's own model: get_detr_model() output (autograd graph alive)
    -> detr_tf/loss/loss.py get_losses (Hungarian matching through the real scipy) -> torch.autograd.grad w.r.t. every
    variable the reference marks trainable (stand-in for tf.GradientTape, training.py:17-23; dropout is the identity: TF's
    random stream cannot be reproduced, the gradient arithmetic is what is pinned)."""
    from detr_tf.loss import hungarian_matching as ref_hm
    from detr_tf.loss import loss as ref_loss
    seed, B, H, W, ne, nd, n_t = TRAIN
    P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd)
    img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
    t_bbox, t_class = O.synthetic_targets(B, n=n_t, seed=seed)
    shim.set_weights({k: v.float() for k, v in P.items()}, autograd=True)
    x = shim.set_input(img)
    cfg = Cfg()
    cfg.background_class = 91
    model = ref_detr.get_detr_model(cfg, include_top=True, num_decoder_layers=nd, num_encoder_layers=ne)
    res = model(x)
    captured = []
    orig = ref_hm.np_tf_linear_sum_assignment

    def spy(matrix):
        r = orig(matrix)
        captured.append((np.array(matrix, copy=True), np.asarray(r[0]).copy(), np.asarray(r[1]).copy()))
        return r
    ref_hm.np_tf_linear_sum_assignment = spy
    try:
        total, losses = ref_loss.get_losses(res, shim.as_tf(t_bbox.float()), shim.as_tf(t_class), cfg)
    finally:
        ref_hm.np_tf_linear_sum_assignment = orig
    names = sorted(n for n, _, t in shim.STATE["created"] if t)
    grads = torch.autograd.grad(total, [shim.STATE["variables"][n] for n in names], allow_unused=True)
    # assignments in call order: main output (all images), then aux 0, 1, ... -> [L, B, Q], layer order aux.., main last
    Q = 100
    match = -np.ones((nd, B, Q), np.int64)
    assert len(captured) == nd * B
    for i, (_, rows, cols) in enumerate(captured):
        layer = (nd - 1) if i < B else (i // B - 1)
        match[layer, i % B, rows] = cols
    out = {"train_meta": np.array(TRAIN), "train_t_bbox": t_bbox.numpy(), "train_t_class": t_class.numpy(),
           "train_total": np.float32(float(total)), "train_match": match,
           "train_loss_keys": np.array(sorted(losses)), "train_loss_values": np.array([float(losses[k]) for k in sorted(losses)], np.float32),
           "train_names": np.array(names)}
    norms, projs = [], []
    for i, (n, g) in enumerate(zip(names, grads)):
        g = torch.zeros_like(P[n]).float() if g is None else g.detach().float()
        r = torch.randn(g.shape, generator=torch.Generator().manual_seed(1000 + i))
        norms.append(float(g.norm()))
        projs.append(float((g * r).sum()))
        if g.numel() <= 2048 or n in ("backbone/conv1/kernel", "class_embed/kernel"):
            out["train_grad/" + n] = g.numpy()
    out["train_grad_norms"] = np.array(norms, np.float64)
    out["train_grad_projs"] = np.array(projs, np.float64)
    print("train: total", float(total), "variables", len(names), "unused", sum(g is None for g in grads),
          "full gradients stored", sum(k.startswith("train_grad/") for k in out))
    return out


def accumulate_case():
    """optimizers.py:137-163 (aggregate_grad_and_apply) executed as is, with a recording stand-in for the Keras optimizer:
    which steps zero the accumulator, which apply, and what is applied (the SUM of the micro-step gradients, not the mean),
    for target_batch // batch_size = 3 over 7 steps, for no accumulation, and for a group whose train_<group> flag is off."""
    from detr_tf import optimizers as ref_opt

    class Recorder:
        def __init__(self):
            self.calls = []

        def apply_gradients(self, pairs):
            self.calls.append([float(g.sum()) for g, _ in pairs])

    class Cfg2:
        batch_size = 2
        train_backbone = True
        train_transformers = False

    out = {}
    for label, target_batch in (("acc3", 6), ("none", None)):
        cfg = Cfg2()
        cfg.target_batch = target_batch
        rec = {"backbone": Recorder(), "transformers": Recorder()}
        variables = [torch.zeros(3), torch.zeros(2, 2)]
        optimizers = {"backbone_optimizer": rec["backbone"], "backbone_variables": variables,
                      "transformers_optimizer": rec["transformers"], "transformers_variables": variables}
        applied_at = []
        for step in range(7):
            grads = [torch.full((3,), float(step + 1)), torch.full((2, 2), 10.0 * (step + 1))]
            for name in ("backbone", "transformers"):
                before = len(rec[name].calls)
                ref_opt.aggregate_grad_and_apply(name, optimizers, grads, step, cfg)
                if len(rec[name].calls) > before:
                    applied_at.append((step, name))
        assert not rec["transformers"].calls                                  # train_transformers = False: never applied
        out[f"accum_{label}_steps"] = np.array([s for s, _ in applied_at])
        out[f"accum_{label}_sums"] = np.array(rec["backbone"].calls, np.float64)     # [n_applies, 2]: sum over each gradient tensor
        print("accumulate", label, "applied at", out[f"accum_{label}_steps"].tolist(), out[f"accum_{label}_sums"].tolist())
    return out


def resnet101_case():
    """BASELINE configs[3] uses a ResNet-101 backbone: networks/resnet_backbone.py:52-66 (ResNet101Backbone, 23 bottlenecks in
    layer3) executed on the shim.  (The reference's DETR class always instantiates ResNet50Backbone, detr.py:31; the deeper
    backbone is selected in the product by get_detr_model(..., backbone='resnet101').)"""
    from detr_tf.networks import resnet_backbone as ref_rb
    seed = 31
    P = O.init_params(seed=seed, backbone="resnet101", num_encoder_layers=1, num_decoder_layers=1)
    img = torch.randn(1, 64, 96, 3, generator=torch.Generator().manual_seed(seed))
    shim.set_weights({k: v.float() for k, v in P.items() if k.startswith("backbone/")})
    with torch.no_grad():
        net = ref_rb.ResNet101Backbone(name="backbone")
        feat = net(shim.as_tf(img))
    created = {n for n, _, _ in shim.STATE["created"]}
    assert created == {k for k in P if k.startswith("backbone/")}, len(created)
    print("resnet101: feat", tuple(feat.shape), "variables", len(created))
    return {"r101_meta": np.array([seed, 1, 64, 96]), "r101_feat": feat.numpy()}


def config_case():
    """training_config.py executed as is: the attribute defaults of TrainingConfig() and the defaults of its argument parser"""
    import json
    from detr_tf import training_config as ref_tc
    c = ref_tc.TrainingConfig()
    attrs = {k: (list(v) if isinstance(v, tuple) else v) for k, v in vars(c).items() if k != "data"}
    attrs = {k: (float(v) if k.endswith("_lr") else v) for k, v in attrs.items()}
    parser = {a.dest: a.default for a in ref_tc.training_config_parser()._actions if a.dest != "help"}
    c.add_nlayers([type("L", (), {"name": "cls_layer"})(), type("L", (), {"name": "pos_layer"})()])
    print("config:", len(attrs), "attributes,", len(parser), "flags")
    return {"config_attrs_json": np.array(json.dumps(attrs, sort_keys=True)), "config_parser_json": np.array(json.dumps(parser, sort_keys=True)),
            "config_nlayers_json": np.array(json.dumps(c.nlayers))}


def main():
    torch.manual_seed(0)
    out = {}
    for case, (seed, B, H, W, ne, nd, nb_class) in CASES.items():
        P = O.init_params(seed=seed, num_encoder_layers=ne, num_decoder_layers=nd, nb_class=nb_class)
        img = torch.randn(B, H, W, 3, generator=torch.Generator().manual_seed(seed))
        shim.set_weights({k: v.float() for k, v in P.items()})
        x = shim.set_input(img)
        with torch.no_grad():
            cfg = Cfg()
            model = ref_detr.get_detr_model(cfg, include_top=nb_class is None, nb_class=nb_class, num_decoder_layers=nd,
                                            num_encoder_layers=ne)
            res = model(x)
        if nb_class is not None:
            assert cfg.nlayers == ["cls_layer", "pos_layer"]
        created = {n: s for n, s, _ in shim.STATE["created"]}
        assert set(created) == set(P), (sorted(set(P) - set(created)), sorted(set(created) - set(P)))
        assert all(tuple(P[n].shape) == created[n] for n in P)
        trainable = sorted(n for n, _, t in shim.STATE["created"] if t)
        acts = shim.STATE["outputs"]
        out[f"{case}_meta"] = np.array([seed, B, H, W, ne, nd, nb_class or 0])
        out[f"{case}_feat"] = acts["backbone"].numpy()                       # [B, h, w, 2048]
        out[f"{case}_pos"] = acts["position_embedding_sine"].numpy()         # [B, h, w, 256]
        out[f"{case}_hs"] = acts["transformer"][0].numpy()                   # [L, B, 100, 256]
        out[f"{case}_memory"] = acts["transformer"][1].numpy()               # [B, h, w, 256]
        out[f"{case}_pred_logits"] = res["pred_logits"].numpy()
        out[f"{case}_pred_boxes"] = res["pred_boxes"].numpy()
        for i, a in enumerate(res["aux"]):
            out[f"{case}_aux{i}_logits"] = a["pred_logits"].numpy()
            out[f"{case}_aux{i}_boxes"] = a["pred_boxes"].numpy()
        out[f"{case}_trainable"] = np.array(trainable)
        print(case, "feat", out[f"{case}_feat"].shape, "hs", out[f"{case}_hs"].shape, "aux", len(res["aux"]),
              "variables", len(created), "trainable", len(trainable))
    out.update(train_case())
    out.update(accumulate_case())
    out.update(resnet101_case())
    out.update(config_case())
    np.savez_compressed(os.path.join(HERE, "model_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "model_golden.npz"), os.path.getsize(os.path.join(HERE, "model_golden.npz")), "bytes")


if __name__ == "__main__":
    main()
