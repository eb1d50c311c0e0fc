#!/bin/bash
# First GPU call of the next round, in order of importance (gpurun --timeout 900 -- 'bash tools/r2_first_call.sh').
# Run tools/r2_build_variants.sh locally first so that build_variants/*.so travel with the snapshot.
#   1. the whole GPU suite (first execution of the filter / voxel-grid / host-call / DLPack tests)
#   2. the bench line, with the host-buffer call timed beside the e2e path
#   3. per-variant event timing + bench (tools/r2_ab_gpu.sh)
#   4. config-3 operators including the new filters
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -rxX --timeout 300 --timeout-method=thread > gpurun_out/r2_pytest.log 2>&1; tail -25 gpurun_out/r2_pytest.log
timeout 200 python bench.py --e2e-host-call > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 900 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
if ls build_variants/*.so >/dev/null 2>&1; then timeout 500 bash tools/r2_ab_gpu.sh > /dev/null 2>&1; cat gpurun_out/r2_ab.txt | cut -c1-400; fi
timeout 300 python tools/bench_ops.py --filters --reps 3 > gpurun_out/r2_ops.json 2> gpurun_out/r2_ops.err; tail -c 1500 gpurun_out/r2_ops.json; tail -3 gpurun_out/r2_ops.err
