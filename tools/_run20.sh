for i in 1 2 3 4 5 6; do CPHB_DEBUG_TIMING=1 timeout 100 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/t$i.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('run $i', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']))"; grep cphb gpurun_out/t$i.err | sed -n '4,6p;9,11p'; done
