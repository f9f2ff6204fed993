#!/bin/bash
# GPU session 9: whole suite, smoke, FIR stage bench + ncu, default bench
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15 ) > gpurun_out/r2_gpu_tests9.log
echo "== tests done" >&2
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 ) > gpurun_out/r2_smoke.log
timeout 600 python tools/fir_stage_bench.py > gpurun_out/r2_fir_stage.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fir_stage_kernel -c 1 -f -o gpurun_out/r2_fir_stage python tools/fir_stage_bench.py > gpurun_out/r2_ncu_fir.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1c.json 2> gpurun_out/r2_bench_n1c.err
echo "== bench done rc=$?" >&2
tail -6 gpurun_out/r2_gpu_tests9.log; cat gpurun_out/r2_smoke.log; grep -E '"config": "c1"|"config": "c3"' gpurun_out/r2_fir_stage.log | cut -c1-330; tail -3 gpurun_out/r2_bench_n1c.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1c.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',json.dumps(d['e2e'])[:600])
print('single',d['single_stream']); print('frontend',json.dumps(d['frontend'])[:700])
print('c4',d['c4']['value'],d['c4']['phase_ms'])
P
