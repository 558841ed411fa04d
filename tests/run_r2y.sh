# round 2, GPU call Y: dropout in the straight-line epilogue (stream kernel on FFN1, persistent fast path), shortcut reorder: tests + bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2y.log
tail -4 gpurun_out/pytest_r2y.log
timeout 600 python bench.py > gpurun_out/bench_r2y.json 2> gpurun_out/bench_r2y.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2y.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2y.csv python tests/profile_step.py > gpurun_out/profile_step_r2y.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_r2y.csv 70 > gpurun_out/launches_r2y_summary.txt; head -12 gpurun_out/launches_r2y_summary.txt
