#!/bin/bash
# bench.py on N GPUs the way the driver launches it: tools/gpu_runN.sh N
cd "$(dirname "$0")/../.."
N=$1
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err
echo "bench rc=$?" >&2
tail -3 gpurun_out/r2_bench_n$N.err; python - <<P
import json
d=json.loads(open('gpurun_out/r2_bench_n$N.json').read().strip().split('\n')[-1])
print('value',d['value'],'e2e',d['e2e']['value'], d['e2e'].get('link_bound_msps'), d['e2e']['matches_device_path'])
for k in ('c4','c4_weak'):
    c=d[k]; print(k, c['value'], c['scaling'], c['stream_samples'], {a:round(b,1) for a,b in c['phase_ms'].items()}, c['tier_s']['frac_gt_1lsb'], [round(p['frac_gt_1lsb'],4) for p in c['tier_s']['per_rank']], c['min_boundary_agreement'], c['kernel']); print('   per rank', c['phase_ms_per_rank'])
P
