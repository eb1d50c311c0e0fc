#!/bin/bash
# Round 2, session 2, GPU call 3: parity of the new default build, then sweeps of the certificate margin (gain, cap) and
# of the helper-block count with per-launch events.
mkdir -p gpurun_out
T=${TAG:-s2c3}
timeout 900 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_geometry.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config4_gicp_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config5_colored_pyramid_2m_vs_oracle tests/test_facade_cpp.py tests/test_gpu_zz_filters.py -m gpu -q --timeout 300 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1; tail -6 gpurun_out/${T}_pytest.log
: > gpurun_out/${T}_sweep.txt
run() {  # label, env...
  echo "=== $*" >> gpurun_out/${T}_sweep.txt
  env "$@" CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 100 python tools/one_registration.py --warm 1 2>&1 | grep -A1 -E "per launch|certificates|tile loops|timeline" | grep -v "^--" | cut -c1-1500 >> gpurun_out/${T}_sweep.txt
  env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']))" >> gpurun_out/${T}_sweep.txt
}
run X=default
for g in 1 0.5 0.3; do for c in 0.5 1 2; do run CPHB_CERT_GAIN=$g CPHB_CERT_CAP=$c; done; done
run CPHB_CERT_GAIN=4 CPHB_CERT_CAP=1
run CPHB_CERT_GAIN=2 CPHB_CERT_CAP=1
for h in 36 72 144; do run CPHB_HELPER_BLOCKS=$h; done
cat gpurun_out/${T}_sweep.txt | grep -E "===|value|tile loops| 0:" | cut -c1-420
