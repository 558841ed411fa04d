# round 2, GPU call AB: interleaved pixel blocks in the weight-gradient kernel (L2 sharing of dY with the data gradient): tests, A/B bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2ab.log
tail -4 gpurun_out/pytest_r2ab.log
for il in 1 0 1 0; do
  DETRB_WGRAD_INTERLEAVE=$il timeout 300 python bench.py --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_r2ab_il$il.json 2> gpurun_out/bench_r2ab_il$il.err
  python - <<PY
import json
d = json.load(open("gpurun_out/bench_r2ab_il$il.json"))
print("interleave=$il", {k: round(d[k], 3) for k in ("value", "ms_per_step")}, round(d["e2e"]["value"], 1), d["e2e"].get("whole_call_img_per_s"), round(d["roofline"]["frac"], 3), {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
done 2>&1 | tee gpurun_out/wgrad_interleave_ab.log
