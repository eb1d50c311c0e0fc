#!/bin/bash
# 2-GPU validation of the round-1 multi-GPU additions (gpurun --gpus 2 --timeout 600 -- 'bash tools/r2_multi_gpu_call.sh'):
# sharded ICP with certificates vs 1 GPU, sharded VoxelDownSample / EstimateNormals vs 1 GPU, config 5 with sharded
# pre-processing.
mkdir -p gpurun_out
P=$((20000 + RANDOM % 20000))
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P tools/multi_gpu_check.py > gpurun_out/r2_mgpu_check.json 2> gpurun_out/r2_mgpu_check.err; tail -c 1500 gpurun_out/r2_mgpu_check.json; tail -3 gpurun_out/r2_mgpu_check.err
for flag in "" "--sharded-prep"; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P + 1)) tools/bench_configs.py --config 5 --points 5000000 --reps 2 $flag > gpurun_out/r2_cfg5$flag.json 2> gpurun_out/r2_cfg5$flag.err; tail -c 1200 gpurun_out/r2_cfg5$flag.json; tail -2 gpurun_out/r2_cfg5$flag.err
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $((P + 2)) bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -c 600 gpurun_out/r2_bench_n2.json
