# round 2, GPU call F: the whole GPU suite in one process (as the driver runs it) + the 2-GPU public-API path
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -40) > gpurun_out/pytest_r2f_all.log
tail -3 gpurun_out/pytest_r2f_all.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_r2f.log
