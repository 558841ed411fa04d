# round 2, GPU call Z (2 GPUs): data-parallel bench + the gloo/nccl DP test, R101 config on one GPU
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2z_dp2.json 2> gpurun_out/bench_r2z_dp2.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2z_dp2.json').read().strip().splitlines()[-1]); print('dp2', d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d['loss_after'])" || tail -5 gpurun_out/bench_r2z_dp2.err
CUDA_VISIBLE_DEVICES=0 timeout 400 python bench.py --steps 20 --warmup 3 --backbone resnet101 --batch 4 --no-cpu-baseline --no-matcher-bench > gpurun_out/bench_r2z_c4.json 2> gpurun_out/bench_r2z_c4.err
python -c "
import json; d=json.loads(open('gpurun_out/bench_r2z_c4.json').read().strip().splitlines()[-1]); print('c4 r101 b4', d['ms_per_step'], d['value'], d['e2e']['value'])" || tail -5 gpurun_out/bench_r2z_c4.err
