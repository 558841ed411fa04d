mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_dp2_$name.json 2> gpurun_out/bench_dp2_$name.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_dp2_$name.json').read().strip().splitlines()[-1]); print('$name', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['loss_after'])" || tail -5 gpurun_out/bench_dp2_$name.err; }
run overlap X=1
run flat DETRB_DP_OVERLAP=0
run overlap2 X=1
