#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c8}
timeout 900 python -m pytest tests/test_gpu_occgrid.py tests/test_facade_cpp.py tests/test_gpu_search.py tests/test_gpu_icp.py -m gpu -q -rfE --timeout 600 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1; tail -25 gpurun_out/${T}_pytest.log
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), 'knn', round(d['knn']['mqueries_per_sec']), d['step_ms'])"; }
b X=1; b X=1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_icp.csv python tools/one_registration.py --warm 0 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/${T}_launches_icp.csv | sed -n 3,12p
