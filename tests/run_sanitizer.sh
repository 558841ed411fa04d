mkdir -p gpurun_out
(timeout 400 compute-sanitizer --tool memcheck --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -30) > gpurun_out/sanitizer_memcheck.log
tail -12 gpurun_out/sanitizer_memcheck.log
