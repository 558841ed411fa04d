"""Mirror of detr_tf/training.py: run_train_step / run_val_step / fit / eval with the reference's signatures,
log keys and print cadence, driving the sm_100a engine."""
import time

import numpy as np
import torch

from .optimizers import GROUPS, aggregate_grad_and_apply, gather_gradient


def _dev(x, dtype, device):
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    return x.to(device=device, dtype=dtype, non_blocking=True)


def _dev_img(x, config, eng):
    """float images -> fp32 (the reference's input); uint8 frames stay uint8 and are normalised on device (extension)"""
    if isinstance(x, np.ndarray):
        x = torch.from_numpy(x)
    if x.dtype == torch.uint8:
        method = getattr(config, "normalized_method", "torch_resnet")
        if getattr(eng, "input_method", None) != method:
            eng.set_input_normalisation(method)
        return x.to(device=eng.device, non_blocking=True)
    return x.to(device=eng.device, dtype=torch.float32, non_blocking=True)


def _forward_loss(model, images, t_bbox, t_class, config, training, loss_scale, with_grad):
    eng = model.engine
    m_outputs = eng.forward(_dev_img(images, config, eng), training=training)
    t_bbox = _dev(t_bbox, torch.float32, eng.device)
    eng.set_targets(t_bbox, _dev(t_class, torch.int64, eng.device))
    eng.set_global_normalisers(t_bbox)          # data parallel: batch-level normalisers of the GLOBAL batch
    eng.loss(int(config.background_class), loss_scale=loss_scale, with_grad=with_grad)
    total, log = eng.loss_dict()
    return m_outputs, total, dict(log)


def run_train_step(model, images, t_bbox, t_class, optimizers, config):
    """training.py:9-25: forward(training=True) -> get_losses -> / gradient_aggregate -> gradients of every group.
    total_loss and the log scalars are snapshots (safe to read after later steps were enqueued).  m_outputs are VIEWS of the
    engine's resident logits / boxes buffers, overwritten by the next step (the reference returns fresh tensors): clone what you
    keep -- in particular inside fit's on_step hook, which runs after the NEXT step has been enqueued."""
    eng = model.engine
    gradient_aggregate = int(config.target_batch // config.batch_size) if config.target_batch is not None else 1
    t_bbox_d = _dev(t_bbox, torch.float32, eng.device)
    eng.stage_inputs(_dev_img(images, config, eng), t_bbox_d, _dev(t_class, torch.int64, eng.device), direct=True)
    eng.set_global_normalisers(t_bbox_d)
    # the reference traces this function once (@tf.function, training.py:9); here the launch sequence of
    # forward + losses + backward + gradient all-reduce is captured once per input shape as a CUDA graph and replayed
    eng.grads_step(int(config.background_class), 1.0 / gradient_aggregate, use_graph=getattr(config, "use_cuda_graph", True))
    m_outputs = eng.outputs()
    total_loss, log = eng.loss_dict(snapshot=True)     # like the reference's fresh tensors: still this step's after the next one ran
    log = dict(log)
    gradient_steps = gather_gradient(model, optimizers, total_loss, None, config, log)
    return m_outputs, total_loss, log, gradient_steps


def run_train_and_apply_step(model, images, t_bbox, t_class, optimizers, config):
    """run_train_step + aggregate_grad_and_apply for every group (training.py:9-25 + 53-54, optimizers.py:137-163 without
    accumulation) as one replayed CUDA graph: the optimizer runs inside the graph, beside the last weight gradient of the
    backward pass.  Used by fit when config.target_batch is None; same results as the two calls it replaces."""
    eng = model.engine
    t_bbox_d = _dev(t_bbox, torch.float32, eng.device)
    eng.stage_inputs(_dev_img(images, config, eng), t_bbox_d, _dev(t_class, torch.int64, eng.device), direct=True)
    eng.set_global_normalisers(t_bbox_d)
    eng.set_lrs(float(config.backbone_lr), float(config.transformers_lr), float(config.nlayers_lr))
    eng.set_enabled(bool(config.train_backbone), bool(config.train_transformers), bool(getattr(config, "train_nlayers", False)))
    eng.fused_step(int(config.background_class), float(config.gradient_norm_clipping))
    m_outputs = eng.outputs()
    total_loss, log = eng.loss_dict(snapshot=True)
    log = dict(log)
    for g in GROUPS:
        log.update({f"{g}_lr": optimizers[f"{g}_optimizer"]._serialize_hyperparameter("learning_rate")})
        if f"{g}_gradients" not in optimizers:
            from .optimizers import _group_views
            optimizers[f"{g}_gradients"] = _group_views(eng, g, eng.grads)
    return m_outputs, total_loss, log


def run_val_step(model, images, t_bbox, t_class, config):
    """training.py:28-32"""
    return _forward_loss(model, images, t_bbox, t_class, config, False, 1.0, False)


class _Prefetcher:
    """Overlaps the host->device copy of batch i+1 with the compute of batch i: batches are copied on a side stream into two
    alternating device staging sets; run_train_step then reads the images where they lie (Engine._stage_images(direct=True): the
    step's opening layout kernel runs from the staging set, no PCIe time and no extra copy on the critical path).  The reference feeds from tf.data, which prefetches the same way."""

    def __init__(self, iterable, device):
        self.it, self.device = iter(iterable), device
        self.stream = torch.cuda.Stream() if device.type == "cuda" else None
        self.slot = 0
        self.bufs = [None, None]
        self.next = self._load()

    @staticmethod
    def _img_dtype(images):
        return torch.uint8 if str(images.dtype).endswith("uint8") else torch.float32

    def _to(self, x, dtype, old):
        if isinstance(x, np.ndarray):
            x = torch.from_numpy(x)
        if x.device == self.device and x.dtype == dtype:
            return x
        if old is None or old.shape != x.shape or old.dtype != dtype:
            old = torch.empty(x.shape, dtype=dtype, device=self.device)
        old.copy_(x, non_blocking=True)
        return old

    def _load(self):
        try:
            images, t_bbox, t_class = next(self.it)
        except StopIteration:
            return None
        if self.stream is None:
            return (_dev(images, self._img_dtype(images), self.device), _dev(t_bbox, torch.float32, self.device),
                    _dev(t_class, torch.int64, self.device), None)
        old = self.bufs[self.slot] or (None, None, None)
        with torch.cuda.stream(self.stream):
            out = (self._to(images, self._img_dtype(images), old[0]), self._to(t_bbox, torch.float32, old[1]),
                   self._to(t_class, torch.int64, old[2]))
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.bufs[self.slot] = out
        self.slot ^= 1
        return out + (ev,)

    def __iter__(self):
        return self

    def __next__(self):
        cur = self.next
        if cur is None:
            raise StopIteration
        if cur[3] is not None:
            main = torch.cuda.current_stream()
            main.wait_event(cur[3])                              # batch i is on the device ...
            done = torch.cuda.Event()
            done.record(main)                                    # everything enqueued so far has consumed the staging set
            self.stream.wait_event(done)                         # that batch i+1 is about to overwrite (two alternate)
        self.next = self._load()                                 # ... and batch i+1 starts moving while batch i computes
        return cur[:3]


def fit(model, train_dt, optimizers, config, epoch_nb, class_names, on_step=None):
    """training.py:35-65 -- one epoch.  `train_dt`: iterable of (images[B,H,W,3] f32, t_bbox[B,100,4] f32,
    t_class[B,100,1] i64) (numpy arrays or torch tensors, host or device).  Extension: `on_step(step, total_loss, log)` is
    called once per step, in order (the reference only prints every 100 steps).  The call for step i is made after step i+1 has
    been enqueued (and after the loop for the last step): a hook that reads the loss back -- a host sync -- then waits for a step
    that is already finishing while the GPU works on the next one, instead of leaving the GPU idle until the host comes back."""
    t = None
    pending = None
    eng = model.engine
    # without gradient accumulation the step and the optimizer are one replayed graph (run_train_and_apply_step); the first step of
    # a process goes through the two reference calls (it sets up every kernel the graph contains)
    fused_ok = (config.target_batch is None and eng.device.type == "cuda" and getattr(config, "use_cuda_graph", True)
                and getattr(config, "fused_optimizer_step", True))
    for epoch_step, (images, t_bbox, t_class) in enumerate(_Prefetcher(train_dt, model.engine.device)):
        if fused_ok and getattr(eng, "_split_step_done", False):
            m_outputs, total_loss, log = run_train_and_apply_step(model, images, t_bbox, t_class, optimizers, config)
            if config.log:
                _train_log_hook(images, t_bbox, t_class, m_outputs, config, config.global_step, class_names, prefix="train/")
        else:
            m_outputs, total_loss, log, gradient_steps = run_train_step(model, images, t_bbox, t_class, optimizers, config)
            if config.log:
                _train_log_hook(images, t_bbox, t_class, m_outputs, config, config.global_step, class_names, prefix="train/")
            for name in gradient_steps:
                aggregate_grad_and_apply(name, optimizers, gradient_steps[name]["gradients"], epoch_step, config)
            eng._split_step_done = True
        model.engine._ensure_weights()      # one refresh of the bf16 weight copies for all groups, enqueued before any host sync
        if on_step is not None:
            if pending is not None:
                on_step(*pending)
            pending = (epoch_step, total_loss, log)
        if epoch_step % 100 == 0:
            t = t if t is not None else time.time()
            elapsed = time.time() - t
            model.engine.check_matcher_status()            # (this block syncs with the host anyway)
            print(f"Epoch: [{epoch_nb}], \t Step: [{epoch_step}], \t ce: [{float(log['label_cost']):.2f}] \t "
                  f"giou : [{float(log['giou_loss']):.2f}] \t l1 : [{float(log['l1_loss']):.2f}] \t time : [{elapsed:.2f}]")
            t = time.time()
        config.global_step += 1
    if pending is not None:
        on_step(*pending)


def eval(model, valid_dt, config, class_name, evaluation_step=200):
    """training.py:68-87"""
    t = None
    for val_step, (images, t_bbox, t_class) in enumerate(valid_dt):
        m_outputs, total_loss, log = run_val_step(model, images, t_bbox, t_class, config)
        if config.log:
            _valid_log_hook(images, t_bbox, t_class, m_outputs, config, val_step, config.global_step, class_name,
                            evaluation_step=evaluation_step, prefix="train/")
        if val_step % 10 == 0:
            t = t if t is not None else time.time()
            elapsed = time.time() - t
            model.engine.check_matcher_status()
            print(f"Validation step: [{val_step}], \t ce: [{float(log['label_cost']):.2f}] \t "
                  f"giou : [{float(log['giou_loss']):.2f}] \t l1 : [{float(log['l1_loss']):.2f}] \t time : [{elapsed:.2f}]")
        if val_step + 1 >= evaluation_step:
            break


def _train_log_hook(*args, **kwargs):
    """wandb image logging (logger/training_logging.py:92-96) is out of scope; the call site is kept as a hook."""


def _valid_log_hook(*args, **kwargs):
    """logger/training_logging.py:99-106 -- out of scope, see _train_log_hook."""
