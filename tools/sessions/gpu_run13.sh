#!/bin/bash
# GPU session 13: seeded c4 after the host-side fixes (vectorised cut targets, batched estimator); sharded tests
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
: > gpurun_out/r2_c4_seeded2.jsonl
for cw in "262144 32768 8192" "262144 32768 4096" "196608 32768 8192" "131072 32768 8192" "262144 24576 8192"; do
  set -- $cw
  timeout 400 python bench.py --mode sharded --stream-samples 8589934592 --chunk $1 --warm $2 --seed-carrier --seed-nfft $3 --steps 3 --warmup 1 --no-e2e --no-cpu 2>gpurun_out/sweep_err.log | tail -1 >> gpurun_out/r2_c4_seeded2.jsonl
  tail -2 gpurun_out/sweep_err.log | grep -i error
done
timeout 400 python bench.py --mode sharded --stream-samples 8589934592 --steps 3 --warmup 1 --no-e2e --no-cpu 2>/dev/null | tail -1 >> gpurun_out/r2_c4_seeded2.jsonl
python - <<'P'
import json
for l in open('gpurun_out/r2_c4_seeded2.jsonl'):
    try: d=json.loads(l)
    except Exception: print('bad', l[:100]); continue
    print(d['config']['workload'][d['config']['workload'].find('chunks of'):][:40], 'GS/s %.1f'%(d['value']/1e3), 'eps %.4f'%d['tier_s']['frac_gt_1lsb'], 'agree %.4f'%d['min_boundary_agreement'], {k:round(v,1) for k,v in d['phase_ms'].items()})
P
timeout 900 python -m pytest tests/test_sharded.py -m gpu -q 2>&1 | tail -3
