"""Developer tool: the device-resident train step in three arrangements -- (A) one graph for everything (capture_train_step, the
bench's `value`), (B) graph for forward + loss + backward, optimizer eager behind it (what training.fit does), (C) like A without
the deferred stem weight gradient.  python tests/time_step_variants.py"""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import detr_tensorflow_b200 as D  # noqa: E402

B, H, W = 8, 800, 1333


def make():
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, seed=0)
    eng = model.engine
    img, tb, tc = bench.synthetic_batch(B, H, W, seed=0)
    eng.forward(img, training=True)
    eng.set_targets(tb, tc)
    eng.set_lrs(1e-5, 1e-4)
    eng.set_enabled(True, True)
    return model, eng


def timeit(step, n=20):
    for _ in range(4):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        step()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


model, eng = make()
a = timeit(eng.capture_train_step(91, 0.1))
print(f"A one graph (fwd, loss, bwd, optimizer with the deferred stem wgrad): {a:.3f} ms/step", flush=True)
del model, eng
torch.cuda.empty_cache()

model, eng = make()


def split():
    eng.grads_step(91, 1.0)
    eng.optimizer_step(0.1)


b = timeit(split)
print(f"B graph (fwd, loss, bwd) + eager optimizer: {b:.3f} ms/step", flush=True)
del model, eng
torch.cuda.empty_cache()

model, eng = make()
orig = eng.backward


def backward_no_defer(*a, **k):
    k["defer_tail"] = False
    return orig(*a, **k)


eng.backward = backward_no_defer
c = timeit(eng.capture_train_step(91, 0.1))
print(f"C one graph, stem weight gradient joined before the optimizer: {c:.3f} ms/step", flush=True)
