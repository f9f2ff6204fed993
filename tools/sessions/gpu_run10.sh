#!/bin/bash
# GPU session 10 (2 GPUs): C multi-GPU tests after the start-gate change, bench --gpus 2 with c4_weak
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_sharded.py -m gpu -q 2>&1 | tail -5 ) > gpurun_out/r2_gpu_tests10.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench_n2b.json 2> gpurun_out/r2_bench_n2b.err
echo "bench rc=$?" >&2
cat gpurun_out/r2_gpu_tests10.log; tail -3 gpurun_out/r2_bench_n2b.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2_bench_n2b.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e'].get('link_bound_msps'))
for k in ('c4','c4_weak'):
    c=d[k]; print(k, c['value'], c['scaling'], c['stream_samples'], c['phase_ms'], c['tier_s']['frac_gt_1lsb'], c['min_boundary_agreement'])
print('frontend', d['frontend']['viterbi_msym_s'], d['frontend']['equals_oracle'])
P
