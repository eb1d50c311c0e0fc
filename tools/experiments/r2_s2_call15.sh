#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-r2g}
timeout 1500 python -m pytest tests -m gpu -q -s -rfE --timeout 600 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1; tail -3 gpurun_out/${T}_pytest.log; grep -c PARITY gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'loop',round(d['loop']['iters_per_sec']),'knn',round(d['knn']['mqueries_per_sec']),'frac',round(d['roofline']['frac'],4))
print('cert_off',round(d['certificates_off']['value']),'c3 voxel',round(d['config3']['voxel']['ms'],3),'knn10m',round(d['config3']['knn_10m']['ms'],3),'c4',round(d['config4']['value']),round(d['config4']['ms_per_registration'],2),d['parity_vs_cpu_baseline']['pose_delta_frobenius'],d['parity_vs_cpu_baseline']['index_mismatches'])"; tail -2 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> /dev/null; cut -c1-120 gpurun_out/${T}_bench_ref.json
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/${T}_events.txt 2>&1; grep -A1 "per launch" gpurun_out/${T}_events.txt | tail -1 | cut -c1-500
b() { env "$@" timeout 100 python bench.py --steps 8 --warmup 3 --no-cpu --no-extras --no-host-call 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$*', 'value', round(d['value']), 'loop', round(d['loop']['iters_per_sec']), 'knn', round(d['knn']['mqueries_per_sec']))"; }
for t in 8 11 18 24; do b CPHB_TRANSPOSE_MAX=$t; done
for c in 1 4 32; do b CPHB_CLAIM_MAX=$c; done
b CPHB_ICP_SEARCH_BLOCKS_PER_SM=4
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_icp.csv python tools/one_registration.py --warm 0 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/${T}_launches_icp.csv 2>/dev/null | head -12
