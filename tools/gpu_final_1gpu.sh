#!/bin/bash
# Final single-GPU evidence of the round: whole GPU suite, both bench arms, per-launch events, launch lists, config-3
# operators, ncu --set full of the two instances of the iteration kernel.  TAG names the output files.
mkdir -p gpurun_out
T=${TAG:-r2f}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1; nproc >> gpurun_out/${T}_smi.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rfE --durations=8 --timeout 600 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1
tail -14 gpurun_out/${T}_pytest.log; grep -c PARITY gpurun_out/${T}_pytest.log
timeout 600 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_bench.json'))
print('value',round(d['value']),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value']),'loop',round(d['loop']['iters_per_sec']),'knn',round(d['knn']['mqueries_per_sec']),'frac',round(d['roofline']['frac'],4))
print('cert_off',round(d['certificates_off']['value']),'c3 voxel',round(d['config3']['voxel']['ms'],3),'knn10m',round(d['config3']['knn_10m']['ms'],3),'build',round(d['config3']['knn_10m']['index_build_ms'],3))
print('c4',round(d['config4']['value']),round(d['config4']['ms_per_registration'],2),'cpu',round(d['cpu_baseline']['value'],2),d['cpu_baseline']['cores'],d['parity_vs_cpu_baseline'])"; tail -2 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> /dev/null; cut -c1-200 gpurun_out/${T}_bench_ref.json
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/${T}_events.txt 2>&1; grep -E "timeline|tile loops" gpurun_out/${T}_events.txt | cut -c1-300
timeout 400 python tools/bench_ops.py --filters --normals 30 --reps 3 > gpurun_out/${T}_ops.json 2> gpurun_out/${T}_ops.err; python -c "
import json; d=json.load(open('gpurun_out/${T}_ops.json')); print('voxel',d['voxel']['ms_median'],'knn',d['knn_vs_downsampled']['ms_median'],'self',d['knn_self']['ms_median'],'normals',d.get('estimate_normals'))"; tail -2 gpurun_out/${T}_ops.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_icp.csv python tools/one_registration.py --warm 0 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/${T}_launches_icp.csv | head -8
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_ops.csv python tools/ops_once.py > /dev/null 2>&1
cap() {  # role launch name
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:icp_iteration_kernelILi2ELi3ELi${1}E -s $((31 + $2)) -c 1 -o gpurun_out/${T}_$3 -f python tools/one_registration.py --warm 1 > gpurun_out/${T}_ncu_$3.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$3.log
}
cap 1 20 cert_l20
cap 0 3 search_l3
cap 0 0 search_l0
ls -la gpurun_out | grep ${T} | awk '{print $5, $9}'
