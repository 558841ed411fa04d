mkdir -p gpurun_out
for m in fit devfit direct pinned; do MODE=$m timeout 300 python tests/repro_fit_race2.py 2>&1 | grep -v "^Epoch" | tail -1; done | tee gpurun_out/repro_fit_race2_r2r.log
