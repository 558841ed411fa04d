# round 2, GPU call C: accumulation probe, tcgen05 attention forward (kernel tests + engine tests), fixed API / parity tests, bench
mkdir -p gpurun_out
python tests/probe_tc_accumulation.py > gpurun_out/probe_tc_accumulation.log 2>&1; cat gpurun_out/probe_tc_accumulation.log
for f in tests/test_kernels_gpu.py tests/test_engine_gpu.py tests/test_gemm_tc_gpu.py; do
  n=$(basename $f .py)
  (timeout 600 python -m pytest $f -m gpu -q -rA -p no:cacheprovider 2>&1 | tail -150) > gpurun_out/pytest_r2c_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_r2c_$n.log)"
done
for id in tests/test_api_gpu.py::test_gradient_accumulation_target_batch_on_device tests/test_api_gpu.py::test_checkpoint_roundtrip_and_resume_on_device tests/test_parity_gpu.py::test_parity_forward_and_assignment_c2_800x1333; do
  n=$(echo $id | sed 's/[^A-Za-z0-9_.-]/_/g')
  (timeout 600 python -m pytest "$id" -m gpu -q -rA -s -p no:cacheprovider 2>&1 | tail -80) > gpurun_out/pytest_r2c_$n.log
  echo "$id: $(tail -1 gpurun_out/pytest_r2c_$n.log)"
done
(timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err); tail -c 1200 gpurun_out/bench_r2c.json; tail -3 gpurun_out/bench_r2c.err
