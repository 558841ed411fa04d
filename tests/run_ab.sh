mkdir -p gpurun_out
for m in 0 1 2 0 1 2; do
DETRB_K256_TCP=$m timeout 300 python tests/time_step_env.py 3 2>&1 | tail -1
done | tee gpurun_out/ab_k256.log
