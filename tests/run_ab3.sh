mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -5) > gpurun_out/pytest_h.log; tail -2 gpurun_out/pytest_h.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench"
run() { name=$1; shift; (env "$@" timeout 200 $B 2>gpurun_out/ab3_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],3), round(d['value'],1), round(d['e2e']['value'],1), d.get('loss_after'), round(d['roofline_hbm']['us_per_launch'],1))") >> gpurun_out/ab3.log 2>&1; }
: > gpurun_out/ab3.log
run hints X=1
run nohints DETRB_L2_HINTS=0
run hints2 X=1
run nohints2 DETRB_L2_HINTS=0
cat gpurun_out/ab3.log
