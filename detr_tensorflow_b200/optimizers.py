"""Mirror of detr_tf/optimizers.py: three Adam optimizers (backbone / transformers / nlayers) with per-variable
clipnorm, gradient split, accumulation and apply -- backed by one multi-tensor kernel over the engine's flat arenas.
"""
import torch

from . import ops

GROUPS = ("backbone", "transformers", "nlayers")


class AdamHandle:
    """Stands in for tf.keras.optimizers.Adam(learning_rate=callable, clipnorm=...) (optimizers.py:86-88): the state
    (m, v, iteration count) lives in the engine's arenas; this object only names the group."""

    def __init__(self, engine, group, config):
        self.engine, self.group, self.config = engine, group, config

    @property
    def learning_rate(self):
        return float(getattr(self.config, f"{self.group}_lr"))

    @property
    def iterations(self):
        return int(self.engine.steps[GROUPS.index(self.group)])

    def _serialize_hyperparameter(self, name):      # used by the reference's log (optimizers.py:129-131)
        assert name == "learning_rate"
        return self.learning_rate


def _group_views(engine, group, arena):
    return [arena[o:o + n] for (_, o, n, g, _) in engine.vars if g == group]


def setup_optimizers(model, config):
    """optimizers.py:67-107.  Returns the same dict keys: <g>_optimizer, <g>_variables (later <g>_gradients)."""
    eng = model.engine
    out = {}
    for g in GROUPS:
        out[f"{g}_optimizer"] = AdamHandle(eng, g, config)
        out[f"{g}_variables"] = _group_views(eng, g, eng.params)
    out["_engine"] = eng
    return out


def gather_gradient(model, optimizers, total_loss, tape, config, log):
    """optimizers.py:110-133: split the (already computed) gradients per group and add the lrs to the log."""
    eng = model.engine
    steps = {g: {"gradients": _group_views(eng, g, eng.grads)} for g in GROUPS}
    for g in GROUPS:
        log.update({f"{g}_lr": optimizers[f"{g}_optimizer"]._serialize_hyperparameter("learning_rate")})
    return steps


def _sync_hyper(eng, config):
    eng.set_lrs(float(config.backbone_lr), float(config.transformers_lr), float(config.nlayers_lr))


def aggregate_grad_and_apply(name, optimizers, gradients, step, config):
    """optimizers.py:137-163 for one group `name`: zero the accumulator at step % k == 0, accumulate, apply at
    (step+1) % k == 0 (clipnorm acts on the accumulated gradient, inside apply)."""
    eng = optimizers["_engine"]
    k = None
    if config.target_batch is not None:
        k = int(config.target_batch // config.batch_size)
    if not bool(getattr(config, f"train_{name}")):
        return
    lo_hi = eng.group_range.get(name)
    if lo_hi is None:          # group without variables (e.g. nlayers when include_top=True)
        return
    lo, hi = lo_hi
    src = eng.grads
    if k is not None:
        if eng.acc is None:
            eng.acc = torch.zeros_like(eng.grads)
        # optimizers.py:150-157 on device: one kernel zeroes (at step % k == 0) and adds this micro-step's gradient
        eng.launches += 1
        ops.accumulate(eng.acc[lo:hi], eng.grads[lo:hi], hi - lo, step % k == 0)
        src = eng.acc
        optimizers[f"{name}_gradients"] = _group_views(eng, name, eng.acc)
    else:
        optimizers[f"{name}_gradients"] = gradients
    if k is None or (step + 1) % k == 0:
        _sync_hyper(eng, config)
        eng.apply_group(name, src, float(config.gradient_norm_clipping))
