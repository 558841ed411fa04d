mkdir -p gpurun_out
(timeout 800 compute-sanitizer --tool memcheck --print-limit 3 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k "persistent and (plain or tma_epilogue or persistent_tiles)" 2>&1 | grep -v "^$" | head -80) > gpurun_out/sanitizer_r2t.log
head -70 gpurun_out/sanitizer_r2t.log
echo "---- plain run, blocking launches"
(CUDA_LAUNCH_BLOCKING=1 timeout 300 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k "persistent and (plain or tma_epilogue or persistent_tiles)" 2>&1 | tail -30) | tee gpurun_out/blocking_r2t.log
