mkdir -p gpurun_out
timeout 500 python tests/time_layers.py 2>&1 | tail -6 | tee gpurun_out/time_layers_r2x.log
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2x.csv python tests/profile_step.py > gpurun_out/profile_step_r2x.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_r2x.csv 60 > gpurun_out/launches_r2x_summary.txt; head -22 gpurun_out/launches_r2x_summary.txt
