# round 2, GPU call E: isolated timings + ncu full captures (with source) of the three tcgen05 attention kernels
mkdir -p gpurun_out
python tests/profile_attn.py 1050 6 > gpurun_out/attn_iso_r2e.log 2>&1; cat gpurun_out/attn_iso_r2e.log
python - <<'PY' > gpurun_out/attn_iso_old_r2e.log 2>&1
import sys, subprocess
sys.path.insert(0, ".")
from detr_tensorflow_b200 import ops
ops.set_tc_attn(0)
sys.argv = ["profile_attn.py", "1050", "6"]
exec(open("tests/profile_attn.py").read())
PY
cat gpurun_out/attn_iso_old_r2e.log
for k in attn_fwd_tc_kernel attn_bwd_dq_tc_kernel attn_bwd_dkv_tc_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 1 --launch-count 1 -f -o gpurun_out/r02_$k python tests/profile_attn.py 1050 3 > gpurun_out/ncu_$k.log 2>&1
done
ls -la gpurun_out/r02_attn*.ncu-rep
