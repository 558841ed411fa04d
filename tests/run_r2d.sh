# round 2, GPU call D: tcgen05 attention backward (kernel + engine tests), re-run of the adjusted parity / API tests, bench
mkdir -p gpurun_out
for f in tests/test_kernels_gpu.py tests/test_engine_gpu.py; do
  n=$(basename $f .py)
  (timeout 600 python -m pytest $f -m gpu -q -rA -p no:cacheprovider 2>&1 | tail -200) > gpurun_out/pytest_r2d_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_r2d_$n.log)"
done
for id in tests/test_api_gpu.py::test_gradient_accumulation_target_batch_on_device tests/test_parity_gpu.py::test_parity_train_step_vs_reference_code_golden tests/test_parity_gpu.py::test_parity_gradients_vs_oracle_full_model; do
  n=$(echo $id | sed 's/[^A-Za-z0-9_.-]/_/g')
  (timeout 600 python -m pytest "$id" -m gpu -q -rA -s -p no:cacheprovider 2>&1 | tail -80) > gpurun_out/pytest_r2d_$n.log
  echo "$id: $(tail -1 gpurun_out/pytest_r2d_$n.log)"
done
(timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err); tail -c 900 gpurun_out/bench_r2d.json; tail -3 gpurun_out/bench_r2d.err
