for t in 0 6 10 14 18 22 32; do CPHB_TRANSPOSE_MAX=$t timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tmax $t', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']))"; done
