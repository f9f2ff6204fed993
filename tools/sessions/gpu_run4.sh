#!/bin/bash
# GPU session 4: whole GPU suite with the new tests, default bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -25 ) > gpurun_out/r2_gpu_tests4.log
echo "== tests done" >&2
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1b.json 2> gpurun_out/r2_bench_n1b.err
echo "== bench done rc=$?" >&2
tail -12 gpurun_out/r2_gpu_tests4.log; tail -3 gpurun_out/r2_bench_n1b.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1b.json').read().strip().split('\n')[-1])
for k in ('value','e2e','value_locked','single_stream','frontend'):
    v=d.get(k)
    print(k, json.dumps(v)[:700])
print('c4', d['c4']['value'], d['c4']['ms_per_step'], d['c4']['tier_s']['frac_gt_1lsb'], d['c4']['phase_ms'])
P
