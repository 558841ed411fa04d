# round 2, GPU call V: launch list of the 615 img/s build + ncu --set full of the streaming and halo kernels (final versions)
mkdir -p gpurun_out
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2v.csv python tests/profile_step.py > gpurun_out/profile_step_r2v.log 2>&1
tail -1 gpurun_out/profile_step_r2v.log
python tests/summarize_launches.py gpurun_out/launches_r2v.csv 40 > gpurun_out/launches_r2v_summary.txt; head -30 gpurun_out/launches_r2v_summary.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:conv3x3_halo --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_conv_halo python tests/profile_conv64_kernel.py > gpurun_out/ncu_conv_halo.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_stream --launch-skip 3 --launch-count 1 -f -o gpurun_out/r02_gemm_stream python tests/profile_hbm_kernel.py > gpurun_out/ncu_gemm_stream.log 2>&1
ls -la gpurun_out/r02_conv_halo.ncu-rep gpurun_out/r02_gemm_stream.ncu-rep
