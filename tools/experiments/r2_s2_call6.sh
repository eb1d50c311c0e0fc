#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c6}
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config4_gicp_1m_vs_oracle -m gpu -q -x --timeout 300 --timeout-method=thread 2>&1 | tail -2
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['step_ms'])"; }
b X=1; b X=1
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 2>&1 | grep -A1 -E "per launch|tile loops|timeline" | grep -v "^--" | cut -c1-1300
CPHB_DEBUG_TIMING=1 CPHB_DEBUG_EVENTS=1 timeout 300 python tools/bench_configs.py --config 4 --reps 2 > gpurun_out/${T}_cfg4_n1.json 2> gpurun_out/${T}_cfg4_n1.err; cat gpurun_out/${T}_cfg4_n1.json; tail -4 gpurun_out/${T}_cfg4_n1.err | cut -c1-1500
