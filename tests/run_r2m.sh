# round 2, GPU call M: halo conv kernel: correctness (base-offset field on / off), timing vs im2col
mkdir -p gpurun_out
for bo in 1 0; do
  echo "== DETRB_HALO_BO=$bo"
  (DETRB_HALO_BO=$bo timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k "halo" 2>&1 | tail -6) | tee gpurun_out/pytest_r2m_bo$bo.log
done
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
for halo in (1, 0):
    ops.set_tc_halo(halo)
    fn = lambda: ops.igemm(x, w, M, C, K, C, K, g, bias=bias, relu=True, C=y, ldc=C, out_bits=ob, ldob=C // 8)
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    print(f"conv3x3 64ch 8x200x334 halo={halo}: {us:.1f} us  {2.0 * M * C * K / us / 1e6:.0f} TF/s  {(2 * M * C * 2 + M * 8) / us / 1e3:.0f} GB/s")
PY
timeout 120 python /tmp/time_conv.py 2>&1 | tail -3 | tee gpurun_out/conv_halo_timing_r2m.log
