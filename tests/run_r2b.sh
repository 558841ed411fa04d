# round 2, GPU call B: every GPU test file in its own process (a device fault in one must not poison the others); the new parity /
# API tests one process per test; full logs
mkdir -p gpurun_out
rm -f gpurun_out/pytest_r2b_*.log
for f in tests/test_engine_gpu.py tests/test_gemm_tc_gpu.py tests/test_kernels_gpu.py tests/test_pipeline_gpu.py tests/test_tma_im2col_gpu.py; do
  n=$(basename $f .py)
  (timeout 600 python -m pytest $f -m gpu -q -rA -p no:cacheprovider 2>&1 | tail -150) > gpurun_out/pytest_r2b_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_r2b_$n.log)"
done
for id in $(python -m pytest tests/test_api_gpu.py tests/test_parity_gpu.py -m gpu --collect-only -q -p no:cacheprovider 2>/dev/null | grep "::"); do
  n=$(echo $id | sed 's/[^A-Za-z0-9_.-]/_/g')
  (timeout 600 python -m pytest "$id" -m gpu -q -rA -s -p no:cacheprovider 2>&1 | tail -120) > gpurun_out/pytest_r2b_$n.log
  echo "$id: $(tail -1 gpurun_out/pytest_r2b_$n.log)"
done
