timeout 200 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench5.err > gpurun_out/bench5.json; python -c "
import json; d=json.load(open('gpurun_out/bench5.json')); print('N1', {k:d[k] for k in ('value','e2e','loop','knn','roofline','cpu_baseline','clocks','gpu_launches')})"
timeout 300 python tools/bench_ops.py 2>&1 | tail -1 | cut -c1-1600 | tee gpurun_out/ops2.json
timeout 300 python tools/bench_ops.py --attrs pnc --reps 3 2>&1 | tail -1 | cut -c1-700
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -s 300 -c 120 --csv --log-file gpurun_out/launches_r1_v4.csv python bench.py --steps 1 --warmup 3 --no-cpu > /dev/null 2>&1
python - <<'PY'
import csv,collections
rows=list(csv.reader(open('gpurun_out/launches_r1_v4.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; kn=h.index('Kernel Name'); mv=h.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[hi+1:]:
    if len(r)<=mv: continue
    name=r[kn].split('(')[0].replace('void ','')[:50]
    try: v=float(r[mv].replace(',',''))
    except: continue
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=v
for k,(c,t) in sorted(agg.items(), key=lambda x:-x[1][1])[:8]: print(f'{t/1e3:9.1f} us {c:4d} x {t/c/1e3:8.2f}  {k}')
PY
