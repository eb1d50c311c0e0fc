python __graft_entry__.py smoke 2>&1 | tail -2
timeout 250 python -m pytest tests -m gpu -q --timeout 120 --timeout-method=thread 2>&1 | tail -3
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; python -c "
import json; d=json.load(open('gpurun_out/bench_final.json')); print('N1', {k:d[k] for k in ('value','ms_per_step','e2e','loop','knn','roofline','cpu_baseline','clocks','gpu_launches')})"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:icp_iteration -s 113 -c 1 -o gpurun_out/prof_icp_r1d -f python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 130 --csv --log-file gpurun_out/launches_r1_final.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-600
ls -la gpurun_out/prof_icp_r1d.ncu-rep gpurun_out/launches_r1_final.csv
