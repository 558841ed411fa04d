mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25) > gpurun_out/pytest_d.log
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-matcher-bench"
run() { name=$1; shift; (env "$@" timeout 200 $B 2>gpurun_out/ab_$name.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$name', round(d['ms_per_step'],3), round(d['value'],1), d.get('loss_after'))") >> gpurun_out/ab.log 2>&1; }
: > gpurun_out/ab.log
run default X=1
run all_off DETRB_ONE_STAGE=0 DETRB_DEEP_SMALL=0 DETRB_R_EARLY=0
run no_one_stage DETRB_ONE_STAGE=0
run no_deep DETRB_DEEP_SMALL=0
run no_early DETRB_R_EARLY=0
run os128 DETRB_ONE_STAGE_BN=128
run os256 DETRB_ONE_STAGE_BN=256
run ob2 DETRB_TCP_OB2=2
run default2 X=1
cat gpurun_out/ab.log; tail -5 gpurun_out/pytest_d.log
