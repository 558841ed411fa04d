# round 2, GPU call Q: is the bimodal step-1 loss older than this session's kernels?  (pre-change tree, same script) + ncu of the halo / stream kernels
mkdir -p gpurun_out
{
echo "== old tree (commit before the bit masks / streaming / halo kernels)"
(cd tests/_oldtree && N=24 timeout 300 python tests/repro_fit_race.py 2>&1 | grep -v "^Epoch" | tail -1)
echo "== new tree"
N=24 timeout 300 python tests/repro_fit_race.py 2>&1 | grep -v "^Epoch" | tail -1
} | tee gpurun_out/repro_fit_race_r2q.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_conv_halo python tests/profile_conv64_kernel.py > gpurun_out/ncu_conv_halo.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_stream --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_gemm_stream python tests/profile_hbm_kernel.py > gpurun_out/ncu_gemm_stream.log 2>&1
ls -la gpurun_out/r02_conv_halo.ncu-rep gpurun_out/r02_gemm_stream.ncu-rep
