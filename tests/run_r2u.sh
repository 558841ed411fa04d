# round 2, GPU call U: elect.sync only in the halo kernel: full GPU suite (x1), gemm tests (x2 more), halo timing, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2u.log
tail -4 gpurun_out/pytest_r2u.log
for i in 1 2; do (timeout 600 python -m pytest tests/test_gemm_tc_gpu.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2); done | tee gpurun_out/pytest_r2u_repeat.log
timeout 600 python bench.py > gpurun_out/bench_r2u.json 2> gpurun_out/bench_r2u.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2u.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
