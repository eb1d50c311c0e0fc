#!/bin/bash
# Round 2, session 2, GPU call 1: whole GPU suite, both bench arms, per-launch events, launch lists and ncu --set full
# captures of the HEAD binary (iteration kernel launches 0 / 3 / 20; config-3 kernels).
mkdir -p gpurun_out
T=${TAG:-s2c1}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${T}_smi.txt 2>&1
nproc > gpurun_out/${T}_nproc.txt
if [ -z "$SKIP_TESTS" ]; then
timeout 1500 python -m pytest tests -m gpu -q -s -rfE --durations=15 --timeout 600 --timeout-method=thread > gpurun_out/${T}_pytest.log 2>&1
tail -25 gpurun_out/${T}_pytest.log; grep PARITY gpurun_out/${T}_pytest.log | cut -c1-700
fi
timeout 400 python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 2500 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; tail -c 900 gpurun_out/${T}_bench_ref.json
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/${T}_events.txt 2>&1; tail -6 gpurun_out/${T}_events.txt | cut -c1-1800
timeout 300 python tools/bench_ops.py --filters --reps 3 > gpurun_out/${T}_ops.json 2> gpurun_out/${T}_ops.err; tail -c 1800 gpurun_out/${T}_ops.json; tail -3 gpurun_out/${T}_ops.err
# launch lists
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_icp.csv python tools/one_registration.py --warm 0 > gpurun_out/${T}_ncu_list.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_ops.csv python tools/ops_once.py > gpurun_out/${T}_ncu_list_ops.log 2>&1
# ncu --set full
for L in 0 3 20; do
  timeout 250 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s $((31 + L)) -c 1 -o gpurun_out/${T}_iter_l$L -f python tools/one_registration.py --warm 1 > gpurun_out/${T}_ncu_l$L.log 2>&1
  tail -1 gpurun_out/${T}_ncu_l$L.log
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'voxel_dense_accum|voxel_dense_write|voxel_dense_count|search1_kernel|kd_refine|hilbert_key|gather_points|leaf_box' -c 10 -o gpurun_out/${T}_ops -f python tools/ops_once.py > gpurun_out/${T}_ncu_ops.log 2>&1
tail -1 gpurun_out/${T}_ncu_ops.log
ls -la gpurun_out | tail -30
