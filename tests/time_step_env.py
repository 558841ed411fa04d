"""Developer tool: ms/step of the one-graph device-resident train step (the bench's `value`) under the environment switches of
the calling shell (DETRB_STREAM_PRIO, DETRB_DEC_FORK, ...).  python tests/time_step_env.py [repeats]"""
import os
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import detr_tensorflow_b200 as D  # noqa: E402

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
res = []
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        step()
    e1.record()
    torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
sw = {k: v for k, v in os.environ.items() if k.startswith("DETRB_")}
print(f"{sw}: ms/step {[round(r, 3) for r in res]}  loss {float(eng.a['total'][0]):.4f}  launches {eng.launches_per_step}", flush=True)
