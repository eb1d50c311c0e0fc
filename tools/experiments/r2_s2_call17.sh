#!/bin/bash
mkdir -p gpurun_out
( time timeout 600 python bench.py > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err ) 2>&1 | grep real
python -c "
import json; d=json.load(open('gpurun_out/r2h_bench.json'))
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'c4',round(d['config4']['value']),d['config4']['step_ms'])
c=d['config5']; print('c5',c['value'],c['ms_per_pyramid'],c['step_ms'],c['stages'],c['final']['pose_error_vs_ground_truth'])"; tail -3 gpurun_out/r2h_bench.err
