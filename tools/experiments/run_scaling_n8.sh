# one 8-GPU call: consistency check, scaling sweep N=1,2,4,8 (p2p), N=8 nccl, configs 4 and 5
set -x
nvidia-smi -L | wc -l
P=29600
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 300 $TR --nproc-per-node 8 --master-port $((P+1)) tools/multi_gpu_check.py 2>&1 | tail -3 | cut -c1-2000 | tee gpurun_out/n8_check.json
timeout 200 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu 2>&1 | tail -1 > gpurun_out/scale_n1.json
for n in 2 4 8; do
  timeout 300 $TR --nproc-per-node $n --master-port $((P+10+n)) bench.py --gpus $n --steps 5 --warmup 3 --comm p2p 2>&1 | tail -1 > gpurun_out/scale_n$n.json
done
timeout 300 $TR --nproc-per-node 8 --master-port $((P+30)) bench.py --gpus 8 --steps 5 --warmup 3 --comm nccl 2>&1 | tail -1 > gpurun_out/scale_n8_nccl.json
for f in gpurun_out/scale_n1.json gpurun_out/scale_n2.json gpurun_out/scale_n4.json gpurun_out/scale_n8.json gpurun_out/scale_n8_nccl.json; do
  python -c "
import json,sys
try:
    d=json.load(open('$f')); print('$f', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'loop', round(d['loop']['iters_per_sec']), d['config']['parallelism'][-8:])
except Exception as e: print('$f', 'ERR', e, open('$f').read()[-400:])"
done
timeout 400 $TR --nproc-per-node 8 --master-port $((P+40)) tools/bench_configs.py --config 4 --reps 3 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/config4_n8.json
timeout 200 python tools/bench_configs.py --config 4 --reps 2 2>&1 | tail -1 | cut -c1-1500 | tee gpurun_out/config4_n1.json
timeout 600 $TR --nproc-per-node 8 --master-port $((P+50)) tools/bench_configs.py --config 5 2>&1 | tail -1 | cut -c1-3000 | tee gpurun_out/config5_n8.json
