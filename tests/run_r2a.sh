# round 2, GPU call A: all GPU tests (new parity-precision + API tests included), smoke, bench (C2 and C4), launch list, ncu captures
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -rA 2>&1 | tail -400) > gpurun_out/pytest_r2a.log
tail -5 gpurun_out/pytest_r2a.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_r2a.log
(timeout 500 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err); tail -c 1500 gpurun_out/bench_r2a.json; tail -5 gpurun_out/bench_r2a.err
(timeout 400 python bench.py --steps 20 --warmup 3 --backbone resnet101 --batch 4 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_r2a_c4.json 2> gpurun_out/bench_r2a_c4.err); tail -c 600 gpurun_out/bench_r2a_c4.json; tail -5 gpurun_out/bench_r2a_c4.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2a.csv python tests/profile_step.py > gpurun_out/profile_step_r2a.log 2>&1
tail -1 gpurun_out/profile_step_r2a.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_wgrad3x3 python tests/profile_wgrad_kernel.py 3x3 > gpurun_out/ncu_wgrad3x3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_tc_kernel --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_wgrad1x1 python tests/profile_wgrad_kernel.py 1x1 > gpurun_out/ncu_wgrad1x1.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_conv64 python tests/profile_conv64_kernel.py > gpurun_out/ncu_conv64.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
