#!/bin/bash
# Round-2 GPU call 1: the whole GPU suite (incl. the new BASELINE-size and FLANN parity tests), bench line, per-launch
# events, A/B of the round-1 build variants, ncu captures of the HEAD binary.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c1_smi.txt 2>&1
nproc > gpurun_out/c1_nproc.txt
timeout 1500 python -m pytest tests -m gpu -q -s -rfE --timeout 600 --timeout-method=thread > gpurun_out/c1_pytest.log 2>&1
tail -15 gpurun_out/c1_pytest.log; grep PARITY gpurun_out/c1_pytest.log | cut -c1-600
timeout 300 python bench.py --e2e-host-call > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 1500 gpurun_out/c1_bench.json; tail -3 gpurun_out/c1_bench.err
CPHB_DEBUG_EVENTS=1 CPHB_DEBUG_CERT=1 timeout 120 python tools/one_registration.py --warm 1 > gpurun_out/c1_events.txt 2>&1; tail -6 gpurun_out/c1_events.txt | cut -c1-1500
timeout 700 bash tools/r2_ab_gpu.sh > /dev/null 2>&1
# ncu: launch list of one registration, then --set full of launch 0 / 3 / 20 of the iteration kernel and one reduce kernel
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/c1_launches.csv python tools/one_registration.py --warm 0 > gpurun_out/c1_ncu_list.log 2>&1
for L in 0 3 20; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s $((31 + L)) -c 1 -o gpurun_out/c1_iter_l$L -f python tools/one_registration.py --warm 1 > gpurun_out/c1_ncu_l$L.log 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:icp_reduce_kernel -s 51 -c 1 -o gpurun_out/c1_reduce_l20 -f python tools/one_registration.py --warm 1 > gpurun_out/c1_ncu_red.log 2>&1
ls -la gpurun_out | tail -20
