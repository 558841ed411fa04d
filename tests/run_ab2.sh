mkdir -p gpurun_out
for s in "534400 256 64 r" "534400 256 64 rm" "534400 64 64 -" "534400 64 256 -" "133600 512 128 r" "133600 512 128 rm" "133600 128 512 -"; do timeout 60 python tests/time_gemm.py $s 2>&1 | tail -1; done | tee gpurun_out/ab2_gemm.log
(timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_f.json 2> gpurun_out/bench_f.err); python -c "
import json; d=json.loads(open('gpurun_out/bench_f.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss_after'])"
