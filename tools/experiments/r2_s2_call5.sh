#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c5}
timeout 1500 python -m pytest tests -m gpu -q -rfE --timeout 600 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1; tail -5 gpurun_out/${T}_pytest.log
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'e2e', round(d['e2e']['value']), 'knn', round(d['knn']['mqueries_per_sec']), d['step_ms'])"; }
b X=1; b X=1; b CPHB_HILBERT_LEVELS=10; b CPHB_HILBERT_LEVELS=9; b CPHB_HILBERT_LEVELS=7
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 2>&1 | grep -A1 -E "per launch|tile loops|timeline" | grep -v "^--" | cut -c1-1300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_icp.csv python tools/one_registration.py --warm 0 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/${T}_launches_icp.csv | head -30
timeout 300 python tools/bench_ops.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('voxel', d['voxel']['ms_median'], 'knn', d['knn_vs_downsampled']['ms_median'], 'self', d['knn_self']['ms_median'])"
CPHB_HILBERT_LEVELS=10 timeout 300 python tools/bench_ops.py --reps 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('L10: voxel', d['voxel']['ms_median'], 'knn', d['knn_vs_downsampled']['ms_median'], 'self', d['knn_self']['ms_median'])"
