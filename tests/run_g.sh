mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/pytest_g.log
tail -3 gpurun_out/pytest_g.log
for i in 1 2; do (timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_g$i.json 2> gpurun_out/bench_g$i.err); python -c "
import json; d=json.loads(open('gpurun_out/bench_g$i.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], 8000/d['e2e']['value'], d['loss_after'])"; done
