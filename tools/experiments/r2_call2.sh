#!/bin/bash
# Round-2 GPU call 2: new iteration kernel (index-order target attributes, certified regime with lane accumulators + fused tail)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py tests/test_gpu_baseline_sizes.py::test_config2_p2plane_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config4_gicp_1m_vs_oracle tests/test_gpu_baseline_sizes.py::test_config5_colored_pyramid_2m_vs_oracle tests/test_facade_cpp.py -m gpu -q -x --timeout 300 --timeout-method=thread > gpurun_out/c2_pytest.log 2>&1; tail -12 gpurun_out/c2_pytest.log
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/c2_events.txt 2>&1; tail -6 gpurun_out/c2_events.txt | cut -c1-1400
timeout 600 bash tools/r2_ab_gpu.sh > /dev/null 2>&1; cut -c1-420 gpurun_out/r2_ab.txt
