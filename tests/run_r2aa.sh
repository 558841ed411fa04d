# round 2, GPU call AA: validation after the probe split / e2e timing change: full GPU suite, smoke, bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2aa.log
tail -4 gpurun_out/pytest_r2aa.log
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) | tee gpurun_out/smoke_r2aa.log
timeout 600 python bench.py > gpurun_out/bench_r2aa.json 2> gpurun_out/bench_r2aa.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2aa.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["e2e"]["passes_img_per_s"], d["roofline"]["frac"], d["cpu_baseline"], d["matcher"])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 | cut -c1-400
