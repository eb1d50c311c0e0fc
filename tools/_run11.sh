set -x
nvidia-smi -L | head -3
timeout 300 python tools/bench_configs.py --config 4 --points 300000 --reps 2 2>&1 | tail -2 | cut -c1-1200
timeout 300 python tools/bench_configs.py --config 5 --points 400000 2>&1 | tail -2 | cut -c1-2500
BENCH_DEBUG=1 timeout 200 python bench.py --steps 3 --warmup 3 --no-cpu 2>gpurun_out/dbg.err | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N1', {k:d[k] for k in ('value','e2e','loop')})"
tail -8 gpurun_out/dbg.err
