"""Developer tool: repeat the 3-step fit of test_uint8_frames_through_model_and_fit and report the spread of the losses.
env: HALO / STREAM / PERSIST / PDL / GRAPH / OVERLAP = 0|1 toggles, N = repetitions"""
import os
import sys

import torch

sys.path.insert(0, ".")
import detr_tensorflow_b200 as D  # noqa: E402
from detr_tensorflow_b200 import _lib, ops  # noqa: E402
from oracle import detr_oracle as O  # noqa: E402

env = lambda k, d: int(os.environ.get(k, d))
ops.set_tc_halo(env("HALO", 1))
ops.set_tc_stream(env("STREAM", 1))
ops.set_tc_persistent(env("PERSIST", 1))
_lib.lib().detrb_set_pdl(env("PDL", 1))
P = O.init_params(seed=4, num_encoder_layers=1, num_decoder_layers=2)
cfg = D.TrainingConfig()
cfg.background_class, cfg.batch_size, cfg.target_batch = 91, 2, None
cfg.train_backbone, cfg.train_transformers = True, True
cfg.use_cuda_graph = bool(env("GRAPH", 1))
u8 = torch.randint(0, 256, (2, 96, 128, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(4))
f32 = torch.from_numpy(O.normalized_images(u8.numpy(), "torch_resnet"))
tb, tc = O.synthetic_targets(2, n=4, seed=4)
seen_all = {}
for rep in range(env("N", 12)):
    imgs = f32 if rep % 2 else u8
    model = D.get_detr_model(cfg, include_top=True, params=P, dropout=0.0, num_encoder_layers=1, num_decoder_layers=2)
    model.engine.overlap_wgrad = bool(env("OVERLAP", 1))
    opt = D.setup_optimizers(model, cfg)
    seen = []
    D.training.fit(model, [(imgs, tb, tc)] * 3, opt, cfg, 0, None, on_step=lambda s, t, l: seen.append(float(t)))
    key = tuple(round(v, 3) for v in seen)
    seen_all[key] = seen_all.get(key, 0) + 1
print({k: os.environ.get(k) for k in ("HALO", "STREAM", "PERSIST", "PDL", "GRAPH", "OVERLAP")}, seen_all)
