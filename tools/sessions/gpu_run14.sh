#!/bin/bash
# GPU session 14: whole suite, smoke, default bench, reference arm
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) > gpurun_out/r2_gpu_tests14.log
( timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 ) > gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_n1d.json 2> gpurun_out/r2_bench_n1d.err
echo "== bench rc=$?" >&2
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_ref.json 2> gpurun_out/r2_bench_ref.err
echo "== ref rc=$?" >&2
tail -4 gpurun_out/r2_gpu_tests14.log; cat gpurun_out/r2_smoke.log; tail -3 gpurun_out/r2_bench_n1d.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n1d.json').read().strip().split('\n')[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'e2e',d['e2e']['value'],d['e2e']['link_bound_msps'],d['e2e']['matches_device_path'])
print('locked',json.dumps(d['value_locked'])[:300]); print('oracle_check',json.dumps(d['oracle_check'])[:300])
print('single',d['single_stream']); print('frontend',d['frontend']['viterbi_msym_s'],d['frontend']['equals_oracle'])
c=d['c4']; print('c4', c['value'], c['phase_ms'], c['tier_s']['frac_gt_1lsb'], c['min_boundary_agreement']); print(c['config']['workload'])
print('cpu', d.get('cpu_baseline')); print('clocks', d['clocks'])
r=json.loads(open('gpurun_out/r2_bench_ref.json').read().strip().split('\n')[-1]); print('ref', r['value'], r['cpu_baseline'])
P
