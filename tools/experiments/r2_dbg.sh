#!/bin/bash
mkdir -p gpurun_out
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/dbg_a.txt 2>&1; grep "timeline" gpurun_out/dbg_a.txt | cut -c1-700
cp cupoch_b200/lib/libcupoch_b200.so /tmp/default.so
for v in base pdl mb5; do cp build_variants/$v.so cupoch_b200/lib/libcupoch_b200.so; echo "== $v"; timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['step_ms'])"; done
cp /tmp/default.so cupoch_b200/lib/libcupoch_b200.so
