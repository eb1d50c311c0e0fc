#!/bin/bash
# Round 2, session 2, GPU call 2: two-instance iteration kernel (searching instance at higher occupancy, certified
# instance with the reduce folded in) + deeper certified pipeline: parity tests on the default build, then A/B of the
# build variants (per-launch events + short bench each).
mkdir -p gpurun_out
T=${TAG:-s2c2}
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_geometry.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config4_gicp_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config5_colored_pyramid_2m_vs_oracle tests/test_facade_cpp.py -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1; tail -6 gpurun_out/${T}_pytest.log
cp cupoch_b200/lib/libcupoch_b200.so /tmp/default.so
: > gpurun_out/${T}_ab.txt
for so in build_variants/*.so; do
  v=$(basename $so .so)
  cp $so cupoch_b200/lib/libcupoch_b200.so
  {
    echo "=== $v"
    CPHB_DEBUG_EVENTS=1 timeout 100 python tools/one_registration.py --warm 1 2>&1 | grep -A1 "per launch" | tail -1 | cut -c1-1300
    for r in 1 2; do timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['step_ms'])"; done
  } >> gpurun_out/${T}_ab.txt 2>&1
done
cp /tmp/default.so cupoch_b200/lib/libcupoch_b200.so
cat gpurun_out/${T}_ab.txt
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/${T}_events.txt 2>&1; grep -E "timeline" gpurun_out/${T}_events.txt | cut -c1-600
CPHB_DEBUG_EVENTS=1 timeout 120 python tools/one_registration.py --warm 1 --kind gicp > gpurun_out/${T}_events_gicp.txt 2>&1; tail -3 gpurun_out/${T}_events_gicp.txt | cut -c1-900
