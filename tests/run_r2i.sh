# round 2, GPU call I: GEMM tests (bit masks, streaming kernel) + isolated timings: stream vs one-tile, bits vs bf16 masks
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2i.log
tail -5 gpurun_out/pytest_r2i.log
for st in 0 1; do
for s in "534400 256 64 r" "534400 256 64 ro" "534400 256 64 rm" "534400 256 64 rb" "534400 64 64 -" "534400 64 64 o" "534400 64 256 -" "534400 64 256 o" "534400 64 256 m" "534400 64 256 b" "133600 512 128 r" "133600 512 128 ro" "133600 512 128 rm" "133600 512 128 rb" "133600 128 512 m" "133600 128 512 b" "33600 1024 256 rm" "33600 1024 256 rb"; do
  echo -n "STREAM=$st  "; STREAM=$st timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
done; done | tee gpurun_out/stream_gemms_r2i.log
