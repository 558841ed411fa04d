# final validation of a build on one B200: every GPU test in one process (as the driver runs them), smoke, bench, launch list
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_final.log
tail -4 gpurun_out/pytest_final.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_final.log
timeout 600 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_final.json"))
print({k: round(d[k], 3) for k in ("value", "ms_per_step")}, "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 3),
      {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()}, d["clocks"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
timeout 400 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python tests/profile_step.py > gpurun_out/profile_step_final.log 2>&1
python tests/summarize_launches.py gpurun_out/launches_final.csv 70 > gpurun_out/launches_final_summary.txt; head -4 gpurun_out/launches_final_summary.txt
timeout 300 python bench.py --backbone resnet101 --batch 4 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_final_r101.json 2> gpurun_out/bench_final_r101.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final_r101.json')); print('R101 B=4', round(d['value'],1), 'e2e', round(d['e2e']['value'],1))" || tail -3 gpurun_out/bench_final_r101.err
