"""Developer tool: kernel timeline of ONE replay of the one-graph train step (torch.profiler / CUPTI): start, duration, stream and
name of every kernel -> gpurun_out/trace_step.csv.  python tests/trace_step.py"""
import csv
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import detr_tensorflow_b200 as D  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

B, H, W = 8, 800, 1333
cfg = D.TrainingConfig()
cfg.background_class = 91
model = D.get_detr_model(cfg, include_top=True, seed=0)
eng = model.engine
img, tb, tc = bench.synthetic_batch(B, H, W, seed=0)
eng.forward(img, training=True)
eng.set_targets(tb, tc)
eng.set_lrs(1e-5, 1e-4)
eng.set_enabled(True, True)
step = eng.capture_train_step(91, 0.1)
for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step()
    step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/trace_step.csv", "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["start_us", "dur_us", "stream", "name"])
    t0 = ev[0].time_range.start
    for e in ev:
        w.writerow([f"{e.time_range.start - t0:.2f}", f"{e.time_range.end - e.time_range.start:.2f}", getattr(e, "device_index", 0) if not hasattr(e, "stream") else e.stream, e.name[:90]])
print(len(ev), "kernel records; span", (ev[-1].time_range.end - t0) / 1e3, "ms")
