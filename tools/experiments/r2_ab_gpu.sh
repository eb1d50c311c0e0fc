#!/bin/bash
# GPU side of the variant A/B (gpurun -- 'bash tools/r2_ab_gpu.sh'): for every build_variants/*.so run one
# registration with per-launch event timing and a short bench (RUN_TESTS=1 adds the ICP parity tests); the
# default library is restored afterwards.  Build the variants first with tools/r2_build_variants.sh.
mkdir -p gpurun_out
cp cupoch_b200/lib/libcupoch_b200.so /tmp/default.so
: > gpurun_out/r2_ab.txt
for so in build_variants/*.so; do
  v=$(basename $so .so)
  cp $so cupoch_b200/lib/libcupoch_b200.so
  {
    echo "=== $v"
    if [ -n "$RUN_TESTS" ]; then
      timeout 120 python -m pytest tests/test_gpu_icp.py tests/test_gpu_zz_certificates.py -m gpu -x -q --timeout 100 --timeout-method=thread 2>&1 | tail -1
    fi
    CPHB_DEBUG_EVENTS=1 timeout 60 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep -A1 "per launch" | tail -1 | cut -c1-900
    timeout 60 python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['step_ms'])"
  } >> gpurun_out/r2_ab.txt 2>&1
done
cp /tmp/default.so cupoch_b200/lib/libcupoch_b200.so
cat gpurun_out/r2_ab.txt
