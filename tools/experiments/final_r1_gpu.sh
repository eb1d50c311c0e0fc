#!/bin/bash
# last GPU call of round 1: validate the default build, record the bench line, then A/B the schedule toggles
mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q --timeout 120 --timeout-method=thread > gpurun_out/final_pytest.log 2>&1; tail -2 gpurun_out/final_pytest.log
timeout 150 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
CPHB_DEBUG_EVENTS=1 timeout 60 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep -A1 "per launch" | tail -2 > gpurun_out/final_events_default.txt
CPHB_STATIC_SCHED=1 CPHB_DEBUG_EVENTS=1 timeout 60 python bench.py --steps 1 --warmup 1 --no-cpu 2>&1 >/dev/null | grep -A1 "per launch" | tail -2 > gpurun_out/final_events_static.txt
run() { echo "== $*"; env "$@" timeout 60 python bench.py --steps 4 --warmup 3 --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['loop']['iters_per_sec']), round(d['e2e']['value']), d['step_ms'])"; }
{
run CPHB_NOOP=1
run CPHB_STATIC_SCHED=1
run CPHB_RETILE_MASK=0x2
run CPHB_RETILE_MASK=0x412
run CPHB_STATIC_SCHED=1 CPHB_RETILE_MASK=0x2
run CPHB_CLAIM_MAX=1
} > gpurun_out/final_sweep.txt 2>&1
cat gpurun_out/final_sweep.txt
CPHB_STATIC_SCHED=1 timeout 120 python -m pytest tests/test_gpu_icp.py -m gpu -x -q --timeout 120 --timeout-method=thread > gpurun_out/final_pytest_static.log 2>&1; tail -2 gpurun_out/final_pytest_static.log
