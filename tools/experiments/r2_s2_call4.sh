#!/bin/bash
# ncu --set full of the two instances of the iteration kernel (HEAD): ROLE 1 at launch 20 (certified regime), ROLE 0 at
# launch 3 (searching regime), ROLE 1 at launch 3 (sum + solve after a searching launch); short bench first.
mkdir -p gpurun_out
T=${TAG:-s2c4}
cap() {  # role launch name
  timeout 300 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:icp_iteration_kernelILi2ELi3ELi${1}E -s $((31 + $2)) -c 1 -o gpurun_out/${T}_$3 -f python tools/one_registration.py --warm 1 > gpurun_out/${T}_ncu_$3.log 2>&1
  tail -1 gpurun_out/${T}_ncu_$3.log
}
cap 1 20 cert_l20
cap 0 3 search_l3
cap 1 3 reduce_l3

ls -la gpurun_out | grep ${T}
