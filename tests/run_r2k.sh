# round 2, GPU call K: streaming kernel with the straight-line epilogue: tests + timings
mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider -k "stream or one_bit" 2>&1 | tail -8) > gpurun_out/pytest_r2k.log
tail -3 gpurun_out/pytest_r2k.log
{
for s in "534400 256 64 r" "534400 256 64 ro" "534400 256 64 rb" "534400 64 64 -" "534400 64 64 o" "534400 64 256 -" "534400 64 256 o" "534400 64 256 b" "133600 512 128 r" "133600 512 128 ro" "133600 512 128 rb" "534400 128 256 o"; do
  timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
done
for d in 2 1; do echo -n "diag=$d  "; DETRB_STREAM_DIAG=$d timeout 120 python tests/time_gemm.py 534400 256 64 r 2>&1 | tail -1; done
for cfg in "2 10" "4 6" "3 8"; do set -- $cfg; echo -n "nst=$1 rs=$2  "; DETRB_STREAM_NST=$1 DETRB_STREAM_RS=$2 timeout 120 python tests/time_gemm.py 534400 256 64 r 2>&1 | tail -1; done
for cfg in "2 8" "4 6" "6 4"; do set -- $cfg; echo -n "nst=$1 rs=$2  "; DETRB_STREAM_NST=$1 DETRB_STREAM_RS=$2 timeout 120 python tests/time_gemm.py 133600 512 128 r 2>&1 | tail -1; done
} 2>&1 | tee gpurun_out/stream_r2k.log
