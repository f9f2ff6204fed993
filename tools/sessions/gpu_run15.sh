#!/bin/bash
# GPU session 15: ncu launch lists (gpu__time_duration only) of the round-2 bench command, our kernels only
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
K='regex:demod_|shard_|fe_|fir_stage|cursor_'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_main_step.csv \
  python bench.py --steps 2 --warmup 1 --no-locked --no-single --no-frontend --no-c4 --no-cpu > gpurun_out/ncu_b1.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$K" -c 400 --csv --log-file gpurun_out/r2_launches_whole_bench.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu --c4-samples 1073741824 > gpurun_out/ncu_b2.log 2>&1
python - <<'P'
import csv, collections
for f in ('gpurun_out/r2_launches_main_step.csv','gpurun_out/r2_launches_whole_bench.csv'):
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    agg=collections.OrderedDict()
    for r in rows:
        name=r[4].split('(')[0][:70]; t=float(r[-1].replace(',',''))
        unit=r[-2]
        ms = t/1e6 if unit in ('ns','nsecond') else t/1e3 if unit in ('us','usecond') else t
        a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=ms
    print(f, len(rows), 'launches')
    for k,(n,ms) in agg.items(): print('   %-70s x%-4d %10.3f ms' % (k,n,ms))
P
