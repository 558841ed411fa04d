mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -q -k "sliding or stem" 2>&1 | tail -15) > gpurun_out/pytest_stem.log
cat gpurun_out/pytest_stem.log | tail -8
(timeout 120 python tests/time_stem.py 2>&1 | tail -12) | tee gpurun_out/time_stem.log
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/pytest_e.log
tail -4 gpurun_out/pytest_e.log
(timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_e.json 2> gpurun_out/bench_e.err); python -c "
import json; d=json.loads(open('gpurun_out/bench_e.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss_after'])"
