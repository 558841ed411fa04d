"""Developer tool (not a test): one eager train step (B=8, 800x1333) between cudaProfilerStart/Stop, for
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tests/profile_step.py
The launch list it yields is per-launch, cold-cache and serialised: use it for SHARES of the step, not absolute times."""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import detr_tensorflow_b200 as D  # noqa: E402

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 800, 1333)
cfg = D.TrainingConfig()
cfg.background_class = 91
model = D.get_detr_model(cfg, include_top=True, seed=0)
eng = model.engine
img, tb, tc = bench.synthetic_batch(B, H, W, seed=0)
eng.forward(img, training=True)
eng.set_targets(tb, tc)
eng.set_lrs(1e-5, 1e-4)
eng.set_enabled(True, True)
for _ in range(2):
    eng.train_step(91, 0.1)
torch.cuda.synchronize()
torch.cuda.profiler.start()
eng.train_step(91, 0.1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("loss", float(eng.a["total"][0]), "launches/step", eng.launches)
