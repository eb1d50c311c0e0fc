#!/bin/bash
# ncu --set full of selected launches of the iteration kernel in one config-2 registration: args = launch indices
mkdir -p gpurun_out
for L in "$@"; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:icp_iteration_kernel -s $((31 + L)) -c 1 -o gpurun_out/${TAG:-c3}_iter_l$L -f python tools/one_registration.py --warm 1 > gpurun_out/${TAG:-c3}_ncu_l$L.log 2>&1
  tail -2 gpurun_out/${TAG:-c3}_ncu_l$L.log
done
