#!/bin/bash
N=${N:-2}
mkdir -p gpurun_out
T=${TAG:-s2m4}_n$N
P=$((20000 + RANDOM % 20000))
CPHB_DEBUG_TIMING=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/bench_configs.py --config 4 --reps 3 > gpurun_out/${T}_cfg4.json 2> gpurun_out/${T}_cfg4.err; cat gpurun_out/${T}_cfg4.json; grep "host ms" gpurun_out/${T}_cfg4.err | tail -$((4 * N))
CPHB_DEBUG_EVENTS=1 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P+1)) tools/bench_configs.py --config 4 --reps 1 2>&1 >/dev/null | grep -A1 "per launch" | tail -4 | cut -c1-900
