#!/bin/bash
# N-GPU validation + timing of the sharded paths (gpurun --gpus N): correctness vs 1 GPU (tools/multi_gpu_check.py), the
# bench line at N ranks (with the config-4 sub-record), per-launch events of rank 0.
N=${N:-2}
mkdir -p gpurun_out
T=${TAG:-s2m}_n$N
P=$((20000 + RANDOM % 20000))
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $P tools/multi_gpu_check.py > gpurun_out/${T}_check.json 2> gpurun_out/${T}_check.err; tail -c 2500 gpurun_out/${T}_check.json; tail -3 gpurun_out/${T}_check.err | cut -c1-300
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 1)) bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; tail -c 3000 gpurun_out/${T}_bench.json; tail -3 gpurun_out/${T}_bench.err | cut -c1-300
if [ -n "$EVENTS" ]; then
CPHB_DEBUG_EVENTS=1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((P + 2)) bench.py --gpus $N --steps 1 --warmup 3 --no-extras > /dev/null 2> gpurun_out/${T}_events.err; grep -A1 "per launch" gpurun_out/${T}_events.err | tail -2 | cut -c1-1200
fi
