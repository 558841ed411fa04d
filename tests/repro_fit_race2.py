"""Developer tool: where does the two-valued step-1 loss of the float-image fit path come from?
MODE=fit (training.fit, host tensors) | devfit (training.fit, device tensors: the prefetcher copies nothing) | direct (run_train_step +
aggregate_grad_and_apply in a plain loop, no prefetcher) | pinned (fit with pinned host tensors)"""
import os
import sys

import torch

sys.path.insert(0, ".")
import detr_tensorflow_b200 as D  # noqa: E402
from detr_tensorflow_b200.optimizers import aggregate_grad_and_apply  # noqa: E402
from oracle import detr_oracle as O  # noqa: E402

mode = os.environ.get("MODE", "fit")
P = O.init_params(seed=4, num_encoder_layers=1, num_decoder_layers=2)
cfg = D.TrainingConfig()
cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 2, None
cfg.train_backbone, cfg.train_transformers = True, True
u8 = torch.randint(0, 256, (2, 96, 128, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
f32 = torch.from_numpy(O.normalized_images(u8.numpy(), "torch_resnet"))
tb, tc = O.synthetic_targets(2, n=4, seed=4)
tb, tc = torch.as_tensor(tb), torch.as_tensor(tc)
res = {}
for rep in range(int(os.environ.get("N", 16))):
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=1, num_decoder_layers=2)
    opt = D.setup_optimizers(model, cfg)
    seen = []
    imgs, b, c = f32, tb, tc
    if mode == "devfit":
        imgs, b, c = f32.cuda(), tb.cuda(), tc.cuda()
    if mode == "pinned":
        imgs, b, c = f32.pin_memory(), tb.pin_memory(), tc.pin_memory()
    if mode == "direct":
        for step in range(3):
            _, total, log, gs = D.training.run_train_step(model, imgs, b, c, opt, cfg)
            for name in gs:
                aggregate_grad_and_apply(name, opt, gs[name]["gradients"], step, cfg)
            seen.append(float(total))
    else:
        D.training.fit(model, [(imgs, b, c)] * 3, opt, cfg, 0, None, on_step=lambda s, t, l: seen.append(float(t)))
    key = tuple(round(v, 3) for v in seen)
    res[key] = res.get(key, 0) + 1
print(mode, res)
