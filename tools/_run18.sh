timeout 200 python -m pytest tests -m gpu -q -x --timeout 120 --timeout-method=thread 2>&1 | tail -3
for b in 8 9 10 12; do CPHB_ICP_BLOCKS_PER_SM=$b timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('blocks/SM $b', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']))"; done
