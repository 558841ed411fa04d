"""Developer tool (not a test): graph-replayed train-step time as a function of the encoder / decoder depth -> live cost of one
encoder layer, one decoder layer and the backbone + fixed part.  python tests/time_layers.py [B H W]"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
import detr_tensorflow_b200 as D  # noqa: E402

B, H, W = (int(x) for x in sys.argv[1:4]) if len(sys.argv) >= 4 else (8, 800, 1333)
res = {}
for ne, nd in ((6, 6), (6, 1), (1, 6), (1, 1)):
    cfg = D.TrainingConfig()
    cfg.background_class = 91
    model = D.get_detr_model(cfg, include_top=True, seed=0, num_encoder_layers=ne, num_decoder_layers=nd)
    eng = model.engine
    img, tb, tc = bench.synthetic_batch(B, H, W, seed=0)
    eng.forward(img, training=True)
    eng.set_targets(tb, tc)
    eng.set_lrs(1e-5, 1e-4)
    eng.set_enabled(True, True)
    step = eng.capture_train_step(91, 0.1)
    for _ in range(3):
        step()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    res[(ne, nd)] = e0.elapsed_time(e1) / 10
    print(f"enc {ne} dec {nd}: {res[(ne, nd)]:.3f} ms/step", flush=True)
    del model, eng, step
    torch.cuda.empty_cache()
enc = (res[(6, 6)] - res[(1, 6)]) / 5
dec = (res[(6, 6)] - res[(6, 1)]) / 5
print(f"per encoder layer {enc:.3f} ms, per decoder layer {dec:.3f} ms, backbone + fixed {res[(6, 6)] - 6 * enc - 6 * dec:.3f} ms")
