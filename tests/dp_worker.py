"""worker for test_data_parallel_two_ranks_gloo (CPU, gloo, C ABI emulated)"""
import sys

import torch
import torch.distributed as dist

import cabi_emulator
from oracle import detr_oracle as O

rank, world, outdir = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
cabi_emulator.install()
cabi_emulator.set_act_dtype(torch.float32)
import detr_tensorflow_b200 as D  # noqa: E402

NE, ND = 1, 2
P = O.init_params(seed=1, num_encoder_layers=NE, num_decoder_layers=ND)
img = torch.randn(2, 32, 48, 3, generator=torch.Generator().manual_seed(5))
tb, tc = O.synthetic_targets(2, n=4, seed=5, n_range=(2, 6))
cfg = D.TrainingConfig()
cfg.background_class = 91


def run(images, tbb, tcc, distributed, bucketed=False):
    model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P)
    eng = model.engine
    eng.forward(images, training=False)
    eng.set_targets(tbb, tcc)
    if distributed:
        eng.set_global_normalisers(tbb)
    eng.zero_grads()
    eng.loss(91)
    if bucketed:                                  # the product's N>1 train step: bucket k reduced while bucket k+1 is computed
        works, order = [], []
        eng.backward(boundary=lambda k: (order.append(k), works.append(eng.allreduce_bucket(k))))
        for w in works:
            w.wait()
        assert order == [0, 1, 2], order
        b = eng.grad_buckets()
        assert sorted(b, key=lambda r: r[0])[0][0] == 0 and sum(hi - lo for lo, hi in b) == eng.total      # a partition of the arena
    else:
        eng.backward()
        eng.allreduce_grads()
    total, _ = eng.loss_dict()
    return float(total), eng.export_grads()


if rank == 0:
    total, grads = run(img, tb, tc, False)
    torch.save({"total": total, "grads": grads}, f"{outdir}/single.pt")
dist.init_process_group("gloo", rank=rank, world_size=world)
total, grads = run(img[rank:rank + 1], tb[rank:rank + 1], tc[rank:rank + 1], True)
t = torch.tensor([total])
dist.all_reduce(t)
_, grads_b = run(img[rank:rank + 1], tb[rank:rank + 1], tc[rank:rank + 1], True, bucketed=True)
# the same step through the PUBLIC API (training.run_train_step: stage_inputs(direct=True) -> grads_step with the bucketed
# all-reduces, global normalisers), dropout 0 so that the training-mode forward equals the one above
model = D.get_detr_model(cfg, include_top=True, num_encoder_layers=NE, num_decoder_layers=ND, device="cpu", params=P, dropout=0.0)
cfg.batch_size, cfg.target_batch = 1, None
cfg.train_backbone = cfg.train_transformers = True
opt = D.setup_optimizers(model, cfg)
_, total_api, log_api, _ = D.training.run_train_step(model, img[rank:rank + 1], tb[rank:rank + 1], tc[rank:rank + 1], opt, cfg)
assert model.engine.s2d_staged                   # the batch was read where it lies, not copied into the resident image buffer
grads_api = model.engine.export_grads()
if rank == 0:
    torch.save({"total_global": float(t), "grads": grads, "grads_bucketed": grads_b, "grads_api": grads_api,
                "total_api": float(total_api)}, f"{outdir}/dp_rank0.pt")
dist.destroy_process_group()
