#!/bin/bash
CPHB_DEBUG_TIMING=1 timeout 300 python tools/bench_configs.py --config 4 --reps 6 2>&1 | grep -E "host ms|ms_per_registration" | cut -c1-400
timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); c=d['config4']; print('bench c4', round(c['value']), c['ms_per_registration'], c['step_ms'], round(c['loop_iters_per_sec']))"
