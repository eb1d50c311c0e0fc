set -x
timeout 300 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -4
timeout 300 python tools/bench_configs.py --config 4 --reps 2 2>&1 | tail -1 | cut -c1-900
timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu --points 125000 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N1-125k', {k:d[k] for k in ('value','loop')})"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 200 --csv --log-file gpurun_out/launches_r1_v3.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_ops.csv python tools/bench_ops.py --reps 1 > /dev/null 2>&1
ls -la gpurun_out/*.csv
