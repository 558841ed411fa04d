# round 2, GPU call G: N1 resize test, live section split of the step, HBM-bound GEMM shapes isolated
mkdir -p gpurun_out
(timeout 300 python -m pytest tests/test_pipeline_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -5) | tee gpurun_out/pytest_r2g.log
(timeout 300 python tests/bench_sections.py 8 800 1333 sections 2>&1 | tail -4) | tee gpurun_out/sections_r2g.log
for s in "534400 256 64 r" "534400 256 64 rm" "534400 64 64 -" "534400 64 256 -" "534400 64 256 m" "133600 512 128 r" "133600 512 128 rm" "133600 128 512 -" "133600 128 512 m" "33600 1024 256 r" "33600 1024 256 rm" "33600 256 1024 -" "33600 256 1024 m"; do
  timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
done | tee gpurun_out/hbm_gemms_r2g.log
