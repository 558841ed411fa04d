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


def _forward_loss(model, images, t_bbox, t_class, config, training, loss_scale, with_grad):
    eng = model.engine
    m_outputs = eng.forward(_dev(images, torch.float32, eng.device), training=training)
    t_bbox = _dev(t_bbox, torch.float32, eng.device)
    eng.set_targets(t_bbox, _dev(t_class, torch.int64, eng.device))
    eng.set_global_normalisers(t_bbox)          # data parallel: batch-level normalisers of the GLOBAL batch
    eng.loss(int(config.background_class), loss_scale=loss_scale, with_grad=with_grad)
    total, log = eng.loss_dict()
    return m_outputs, total, dict(log)


def run_train_step(model, images, t_bbox, t_class, optimizers, config):
    """training.py:9-25: forward(training=True) -> get_losses -> / gradient_aggregate -> gradients of every group."""
    eng = model.engine
    gradient_aggregate = int(config.target_batch // config.batch_size) if config.target_batch is not None else 1
    t_bbox_d = _dev(t_bbox, torch.float32, eng.device)
    eng.stage_inputs(_dev(images, torch.float32, eng.device), t_bbox_d, _dev(t_class, torch.int64, eng.device))
    eng.set_global_normalisers(t_bbox_d)
    # the reference traces this function once (@tf.function, training.py:9); here the launch sequence of
    # forward + losses + backward + gradient all-reduce is captured once per input shape as a CUDA graph and replayed
    eng.grads_step(int(config.background_class), 1.0 / gradient_aggregate, use_graph=getattr(config, "use_cuda_graph", True))
    m_outputs = eng.outputs()
    total_loss, log = eng.loss_dict()
    log = dict(log)
    gradient_steps = gather_gradient(model, optimizers, total_loss, None, config, log)
    return m_outputs, total_loss, log, gradient_steps


def run_val_step(model, images, t_bbox, t_class, config):
    """training.py:28-32"""
    return _forward_loss(model, images, t_bbox, t_class, config, False, 1.0, False)


def fit(model, train_dt, optimizers, config, epoch_nb, class_names):
    """training.py:35-65 -- one epoch.  `train_dt`: iterable of (images[B,H,W,3] f32, t_bbox[B,100,4] f32,
    t_class[B,100,1] i64) (numpy arrays or torch tensors, host or device)."""
    t = None
    for epoch_step, (images, t_bbox, t_class) in enumerate(train_dt):
        m_outputs, total_loss, log, gradient_steps = run_train_step(model, images, t_bbox, t_class, optimizers, config)
        if config.log:
            _train_log_hook(images, t_bbox, t_class, m_outputs, config, config.global_step, class_names, prefix="train/")
        for name in gradient_steps:
            aggregate_grad_and_apply(name, optimizers, gradient_steps[name]["gradients"], epoch_step, config)
        if epoch_step % 100 == 0:
            t = t if t is not None else time.time()
            elapsed = time.time() - t
            print(f"Epoch: [{epoch_nb}], \t Step: [{epoch_step}], \t ce: [{float(log['label_cost']):.2f}] \t "
                  f"giou : [{float(log['giou_loss']):.2f}] \t l1 : [{float(log['l1_loss']):.2f}] \t time : [{elapsed:.2f}]")
            t = time.time()
        config.global_step += 1


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
            print(f"Validation step: [{val_step}], \t ce: [{float(log['label_cost']):.2f}] \t "
                  f"giou : [{float(log['giou_loss']):.2f}] \t l1 : [{float(log['l1_loss']):.2f}] \t time : [{elapsed:.2f}]")
        if val_step + 1 >= evaluation_step:
            break


def _train_log_hook(*args, **kwargs):
    """wandb image logging (logger/training_logging.py:92-96) is out of scope; the call site is kept as a hook."""


def _valid_log_hook(*args, **kwargs):
    """logger/training_logging.py:99-106 -- out of scope, see _train_log_hook."""
