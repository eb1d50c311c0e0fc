#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c12}
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config4_gicp_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config5_colored_pyramid_2m_vs_oracle -m gpu -q --timeout 300 --timeout-method=thread 2>&1 | tail -2
cp cupoch_b200/lib/libcupoch_b200.so /tmp/default.so
for so in build_variants/*.so; do
  v=$(basename $so .so); cp $so cupoch_b200/lib/libcupoch_b200.so; echo "=== $v"
  CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 100 python tools/one_registration.py --warm 1 2>&1 | grep -E "timeline|tile loops" | tail -2 | cut -c1-330
  CPHB_DEBUG_EVENTS=1 timeout 100 python tools/one_registration.py --warm 1 --kind gicp 2>&1 | grep -A1 "per launch" | tail -1 | cut -c1-140
  for r in 1 2; do timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), d['step_ms'])"; done
done
cp /tmp/default.so cupoch_b200/lib/libcupoch_b200.so
