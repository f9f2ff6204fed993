#!/bin/bash
# GPU session 12: time-sharded stream with the seeded warm-up (coarse carrier estimate): chunk / warm-up / nfft sweep
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2_c4_seeded_sweep.jsonl
for cw in "262144 32768 16384" "262144 49152 16384" "131072 32768 16384" "113360 32768 16384" "131072 49152 16384" "131072 32768 4096" "262144 32768 65536"; do
  set -- $cw
  timeout 400 python bench.py --mode sharded --stream-samples 8589934592 --chunk $1 --warm $2 --seed-carrier --seed-nfft $3 --steps 2 --warmup 1 --no-e2e --no-cpu 2>gpurun_out/sweep_err.log | tail -1 >> gpurun_out/r2_c4_seeded_sweep.jsonl
  tail -2 gpurun_out/sweep_err.log | grep -i error
done
python - <<'P'
import json
for l in open('gpurun_out/r2_c4_seeded_sweep.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(d['config']['workload'][d['config']['workload'].find('chunks of'):][:40], 'GS/s %.1f'%(d['value']/1e3), 'eps %.4f'%d['tier_s']['frac_gt_1lsb'], 'agree %.4f'%d['min_boundary_agreement'], {k:round(v,1) for k,v in d['phase_ms'].items()})
P
