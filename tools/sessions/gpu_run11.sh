#!/bin/bash
# GPU session 11: chunk / warm-up sweep of the time-sharded stream (throughput vs Tier-S epsilon and boundary agreement)
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2_c4_sweep.jsonl
for cw in "262144 150000" "262144 98304" "262144 65536" "131072 98304" "131072 65536" "131072 49152" "98304 65536"; do
  set -- $cw
  timeout 400 python bench.py --mode sharded --stream-samples 8589934592 --chunk $1 --warm $2 --steps 2 --warmup 1 --no-e2e --no-cpu 2>/dev/null | tail -1 >> gpurun_out/r2_c4_sweep.jsonl
done
python - <<'P'
import json
for l in open('gpurun_out/r2_c4_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(d['config']['workload'][d['config']['workload'].find('chunks of'):][:40], 'GS/s %.1f'%(d['value']/1e3), 'eps %.4f'%d['tier_s']['frac_gt_1lsb'], 'agree %.4f'%d['min_boundary_agreement'], {k:round(v,1) for k,v in d['phase_ms'].items()})
P
