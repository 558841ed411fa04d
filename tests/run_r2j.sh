# round 2, GPU call J: A/B of the one-tile kernel before / after the bit-mask epilogue; streaming-kernel diagnostics
mkdir -p gpurun_out
{
for s in "534400 256 64 r" "534400 64 256 -" "534400 64 64 -" "534400 64 256 m" "133600 512 128 r"; do
  echo -n "old  "; DETRB_SO=tests/_old/libdetrb_old.so timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
  echo -n "new  "; STREAM=0 timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
done
for d in 0 1 2 4 3 7; do echo -n "diag=$d  "; DETRB_STREAM_DIAG=$d timeout 120 python tests/time_gemm.py 534400 256 64 r 2>&1 | tail -1; done
for cfg in "2 10" "4 8" "4 4" "8 4" "2 6"; do set -- $cfg; echo -n "nst=$1 rs=$2  "; DETRB_STREAM_NST=$1 DETRB_STREAM_RS=$2 timeout 120 python tests/time_gemm.py 534400 256 64 r 2>&1 | tail -1; done
for bn in 128 64; do echo -n "bn=$bn  "; DETRB_STREAM_BN=$bn timeout 120 python tests/time_gemm.py 534400 256 64 r 2>&1 | tail -1; done
echo -n "no residual  "; timeout 120 python tests/time_gemm.py 534400 256 64 - 2>&1 | tail -1
echo -n "no residual diag=1  "; DETRB_STREAM_DIAG=1 timeout 120 python tests/time_gemm.py 534400 256 64 - 2>&1 | tail -1
echo -n "nobias  "; timeout 120 python tests/time_gemm.py 534400 256 64 r -1 0 nobias 2>&1 | tail -1
} 2>&1 | tee gpurun_out/stream_diag_r2j.log
