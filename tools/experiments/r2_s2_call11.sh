#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c11}
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), d['step_ms'])"; }
for bl in 4 3 2 1; do b CPHB_ICP_BLOCKS_PER_SM=$bl; env CPHB_ICP_BLOCKS_PER_SM=$bl CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 100 python tools/one_registration.py --warm 1 2>&1 | grep -E "timeline|tile loops" | tail -2 | cut -c1-330; done
b CPHB_ICP_BLOCKS_PER_SM=3 CPHB_HELPER_BLOCKS=55
b CPHB_ICP_BLOCKS_PER_SM=2 CPHB_HELPER_BLOCKS=37
