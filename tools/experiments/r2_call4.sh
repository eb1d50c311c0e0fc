#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/${TAG}_pytest.log 2>&1; tail -3 gpurun_out/${TAG}_pytest.log
CPHB_DEBUG_EVENTS=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/${TAG}_events.txt 2>&1; tail -3 gpurun_out/${TAG}_events.txt | cut -c1-1200
timeout 100 python bench.py --steps 6 --warmup 3 --no-cpu --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['e2e'].get('call'), d['step_ms'])"
bash tools/r2_ncu.sh $NCU
