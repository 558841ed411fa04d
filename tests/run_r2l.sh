# round 2, GPU call L: full GPU suite + bench with the streaming kernel; layer3 shapes on the streaming kernel vs the one-tile kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -15) > gpurun_out/pytest_r2l.log
tail -4 gpurun_out/pytest_r2l.log
{
for st in 0 1; do for s in "33600 1024 256 r" "33600 1024 256 ro" "33600 1024 256 rb" "33600 256 1024 b" "133600 128 512 b"; do
  echo -n "STREAM=$st  "; STREAM=$st timeout 120 python tests/time_gemm.py $s 2>&1 | tail -1
done; done
} 2>&1 | tee gpurun_out/stream_l3_r2l.log
timeout 600 python bench.py > gpurun_out/bench_r2l.json 2> gpurun_out/bench_r2l.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2l.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["e2e"]["value"], d["roofline"]["frac"], {k: round(v["us_per_launch"], 1) for k, v in d["rooflines"].items()})
PY
