#!/bin/bash
# GPU side of the variant A/B (gpurun -- 'bash tools/r2_ab_gpu.sh'): for every build_variants/*.so run the ICP
# parity tests, one registration with per-launch event timing, and a short bench; restore the default library.
mkdir -p gpurun_out
cp cupoch_b200/lib/libcupoch_b200.so /tmp/default.so
for so in build_variants/*.so; do
  v=$(basename $so .so)
  cp $so cupoch_b200/lib/libcupoch_b200.so
  {
    echo "=== $v"
    timeout 120 python -m pytest tests/test_gpu_icp.py -m gpu -x -q --timeout 100 --timeout-method=thread 2>&1 | tail -1
    CPHB_DEBUG_EVENTS=1 timeout 60 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep -A1 "per launch" | tail -1
    timeout 60 python bench.py --steps 6 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), d['step_ms'])"
  } >> gpurun_out/r2_ab.txt 2>&1
done
cp /tmp/default.so cupoch_b200/lib/libcupoch_b200.so
cat gpurun_out/r2_ab.txt
