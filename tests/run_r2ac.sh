# round 2, GPU call AC: fit with the fused step + optimizer graph: tests, bench (1 GPU)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2ac.log
tail -4 gpurun_out/pytest_r2ac.log
timeout 600 python bench.py --no-matcher-bench > gpurun_out/bench_r2ac.json 2> gpurun_out/bench_r2ac.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2ac.json"))
print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, round(d["e2e"]["value"], 1), d["e2e"]["passes_img_per_s"], d["e2e"].get("whole_call_img_per_s"), round(d["roofline"]["frac"], 3), d["roofline"]["us_per_launch"], d["roofline"]["beside_data_gradient_chain"]["us_per_launch"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
tail -3 gpurun_out/bench_r2ac.err
