#!/bin/bash
mkdir -p gpurun_out
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), d['step_ms'])"; }
for m in 0x12 0x2 0x4 0x6 0xA 0x0 0x22; do b CPHB_RETILE_MASK=$m; done
timeout 300 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py -m gpu -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -1
