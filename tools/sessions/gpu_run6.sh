#!/bin/bash
# GPU session 6 (8 GPUs): bench --gpus 4 and --gpus 8 (c4 sub-record over 4 / 8 ranks); short steps
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
for n in 4 8; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 5 --warmup 3 --no-single --no-frontend > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
  echo "== bench$n rc=$?" >&2
done
python - <<'P'
import json
for n in (4,8):
    try:
        d=json.loads(open('gpurun_out/r2_bench_n%d.json'%n).read().strip().split('\n')[-1])
        print(n,'value',d['value'],'e2e',d['e2e']['value'],d['e2e']['link_bound_msps'],'locked',d['value_locked']['value'])
        c=d['c4']; print(n,'c4',c['value'],c['ms_per_step'],c['kernel'],c['chunks_per_rank'],c['phase_ms'],[ (p['rank'],round(p['frac_gt_1lsb'],5)) for p in c['tier_s']['per_rank']], c['min_boundary_agreement'])
    except Exception as e:
        print(n,'ERR',e)
P
tail -3 gpurun_out/r2_bench_n8.err
