# round 2, GPU call H: full GPU suite + bench (1-bit ReLU masks)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2h.log
tail -5 gpurun_out/pytest_r2h.log
timeout 600 python bench.py > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2h.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
