#!/bin/bash
mkdir -p gpurun_out
T=${TAG:-s2c7}
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${T}_launches_gicp5m.csv python tools/one_registration.py --warm 0 --kind gicp --points 5000000 > /dev/null 2>&1
python tools/launch_breakdown.py gpurun_out/${T}_launches_gicp5m.csv 2>&1 | head -40
