P=29700
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for n in 8 4; do
  timeout 200 $TR --nproc-per-node $n --master-port $((P+n)) bench.py --gpus $n --steps 5 --warmup 3 --comm p2p 2>/dev/null | tail -1 > gpurun_out/scale2_n$n.json
  python -c "
import json
try:
    d=json.load(open('gpurun_out/scale2_n$n.json')); print('N$n', 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'loop', round(d['loop']['iters_per_sec']), 'knn', round(d['knn']['mqueries_per_sec']), 'clocks', d['clocks'])
except Exception as e: print('N$n ERR', e)"
done
timeout 300 $TR --nproc-per-node 8 --master-port $((P+40)) tools/bench_configs.py --config 4 --reps 3 2>/dev/null | tail -1 | cut -c1-1200 | tee gpurun_out/config4_n8_fixed.json
