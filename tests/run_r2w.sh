# round 2, GPU call W: BITS template in the one-tile kernel (small-GEMM regression), early accumulator release: tests, timings, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2w.log
tail -4 gpurun_out/pytest_r2w.log
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
{
for d in 0 2; do echo -n "halo diag=$d  "; DETRB_HALO_DIAG=$d timeout 120 python /tmp/time_conv.py 2>&1 | tail -1; done
for s in "534400 256 64 ro" "534400 256 64 rb" "534400 64 256 b" "534400 64 64 o" "133600 512 128 ro" "33600 1024 256 ro" "8400 256 256 r" "8400 2048 256 -" "800 256 256 r" "800 256 2048 r" "8400 512 256 -"; do timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1; done
} 2>&1 | tee gpurun_out/timings_r2w.log
timeout 600 python bench.py > gpurun_out/bench_r2w.json 2> gpurun_out/bench_r2w.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2w.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
