mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/pytest_final.log
tail -3 gpurun_out/pytest_final.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_final.log
(timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err); python -c "
import json; d=json.loads(open('gpurun_out/bench_final.json').read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['value'], d['e2e']['value'], d['loss_after'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['cpu_baseline'])"
(timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final.json 2> gpurun_out/bench_ref_final.err); tail -c 400 gpurun_out/bench_ref_final.json
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tests/profile_step.py > gpurun_out/profile_step_final.log 2>&1
tail -1 gpurun_out/profile_step_final.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tcp_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/top_tensor python tests/profile_top_kernel.py > gpurun_out/ncu_top_tensor.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/top_hbm python tests/profile_hbm_kernel.py > gpurun_out/ncu_top_hbm.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
