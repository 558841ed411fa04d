# round 2, GPU call N: halo conv diagnostics; full GPU suite; bench
mkdir -p gpurun_out
cat > /tmp/time_conv.py <<'PY'
import sys, torch
sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops
B, H, W, C = 8, 200, 334, 64
M, K = B * H * W, 9 * C
x = torch.randn(M, C, device="cuda").to(torch.bfloat16)
w = (torch.randn(C, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
y = torch.empty(M, C, dtype=torch.bfloat16, device="cuda")
ob = torch.empty(M, C // 8, dtype=torch.uint8, device="cuda")
bias = torch.zeros(C, device="cuda")
g = dict(batch=B, IH=H, IW=W, Cin=C, OH=H, OW=W, KH=3, KW=3, stride=1, pad=1, mode=0)
fn = lambda: ops.igemm(x, w, M, C, K, C, K, g, bias=bias, relu=True, C=y, ldc=C, out_bits=ob, ldob=C // 8)
for _ in range(3): fn()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.synchronize(); e0.record()
for _ in range(20): fn()
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) / 20 * 1e3
print(f"conv3x3 64ch 8x200x334: {us:.1f} us  {2.0 * M * C * K / us / 1e6:.0f} TF/s  {(2 * M * C * 2 + M * 8) / us / 1e3:.0f} GB/s")
PY
for d in 0 1 2 4 5 6; do echo -n "diag=$d  "; DETRB_HALO_DIAG=$d timeout 120 python /tmp/time_conv.py 2>&1 | tail -1; done | tee gpurun_out/conv_halo_diag_r2n.log
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2n.log
tail -4 gpurun_out/pytest_r2n.log
timeout 600 python bench.py > gpurun_out/bench_r2n.json 2> gpurun_out/bench_r2n.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2n.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
