# round 2, GPU call AE: stride-2 scatter through the dense workspace: tests, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2ae.log
tail -4 gpurun_out/pytest_r2ae.log
timeout 600 python bench.py --no-matcher-bench --no-cpu-baseline > gpurun_out/bench_r2ae.json 2> gpurun_out/bench_r2ae.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2ae.json"))
print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, round(d["e2e"]["value"], 1), round(d["roofline"]["frac"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
tail -3 gpurun_out/bench_r2ae.err
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2ae.csv python tests/profile_step.py > gpurun_out/profile_step_r2ae.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_r2ae.csv 70 > gpurun_out/launches_r2ae_summary.txt; head -3 gpurun_out/launches_r2ae_summary.txt; grep -n "scatter_s2\|<128, 3" gpurun_out/launches_r2ae_summary.txt | head
