import sys, torch
sys.path.insert(0, ".")
import detr_tensorflow_b200 as D
from bench import synthetic_batch
B,H,W = 8,800,1333
cfg = D.TrainingConfig(); cfg.background_class = 91
model = D.get_detr_model(cfg, include_top=True, seed=0)
eng = model.engine
img, tb, tc = synthetic_batch(B,H,W,0)
eng.forward(img, training=True)
eng.set_targets(tb, tc); eng.set_lrs(1e-5,1e-4); eng.set_enabled(True, True)
eng.train_step(91, 0.1)
torch.cuda.synchronize()
print("loss", float(eng.a["total"][0]))
bad = []
for k, v in eng.a.items():
    if torch.is_tensor(v) and v.is_floating_point():
        n = int((~torch.isfinite(v.float())).sum())
        if n: bad.append((k, n, v.numel()))
print("nonfinite activations/scratch:", bad[:40])
for l in range(6):
    for nm in ("lse",):
        pass
g = eng.grads
print("grads nonfinite", int((~torch.isfinite(g)).sum()), "params nonfinite", int((~torch.isfinite(eng.params)).sum()))
for name, s in list(eng.slots.items()):
    gg = s.grad
    n = int((~torch.isfinite(gg)).sum())
    if n: print("grad slot", name, n, gg.numel())
