# validation of the handle build (ABI 222) on one B200: every GPU test in one process, smoke, the default bench line
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2g.log
tail -4 gpurun_out/pytest_r2g.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_r2g.log
timeout 600 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2g.json"))
print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3),
      {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()}, d["clocks"])
PY
