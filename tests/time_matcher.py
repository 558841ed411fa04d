"""Developer tool: latency of matcher_kernel at the train step's size (48 problems = 6 decoder layers x 8 images, 100 queries x 20
targets) and at BASELINE configs[4] (256 problems).  python tests/time_matcher.py"""
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from detr_tensorflow_b200 import ops  # noqa: E402

Q, C = 100, 92
for P, B in ((48, 8), (256, 256)):
    logits, boxes, tb, tc = bench.c5_inputs(B, Q, C, 20, 1234, layers=max(1, P // B))
    dl = logits.reshape(-1, Q, C)[:P].contiguous().cuda()
    db = boxes.reshape(-1, Q, 4)[:P].contiguous().cuda()
    dtb, dtc = tb.cuda(), tc.cuda()
    out = dict(p=torch.empty(P, Q, dtype=torch.int64, device="cuda"), t=torch.empty(P, Q, dtype=torch.int64, device="cuda"),
               s=torch.empty(P, Q, dtype=torch.uint8, device="cuda"), m=torch.empty(P, Q, dtype=torch.int32, device="cuda"),
               st=torch.empty(P, dtype=torch.int32, device="cuda"))
    fn = lambda: ops.matcher(dl, C, db, dtb, dtc, P, B, Q, C, out["p"], out["t"], out["s"], out["m"], None, out["st"])
    for _ in range(3):
        fn()
    ts = []
    for _ in range(20):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print(f"matcher {P} problems: median {ts[10]:.1f} us  min {ts[0]:.1f} us   checksum {int(out['m'].long().sum())} status {int(out['st'].sum())}", flush=True)
